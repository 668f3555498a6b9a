"""torchrun helper: PSO / GA with the population's fitness evaluation sharded over WORLD_SIZE GPUs must
reproduce the reference trajectories recorded in tests/golden/solver_cases.npz (bit-exact)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
POLICY = dict(turn_penalty_factor=0.3, safety_penalty_factor=0.8, min_safe_distance=1.8,
              diagonal_obstacle_penalty_value=100.0, allow_diagonal_moves=True,
              restrict_diagonal_near_obstacle_policy=True)


def main():
    from maaco_path_planing_b200.ga_solver import GASolver
    from maaco_path_planing_b200.pso import PSOSolver
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = np.load(os.path.join(ROOT, "tests", "golden", "solver_cases.npz"))
    for name in ("fig7_33", "blocks40_33"):
        k = "pso_" + name
        N, K, seed = (int(x) for x in g[k + "_meta"])
        grid = g[k + "_grid"].astype(int)
        s = PSOSolver(grid, num_iterations=K, num_particles=N, num_waypoints_per_particle=5, w=0.7, c1=1.5, c2=1.5,
                      rng_seed=seed, device=local, verbose=False, group=dist.group.WORLD, **POLICY)
        res = s.solve()
        assert np.array_equal(np.array(s.convergence_curve), g[k + "_curve"]), "pso curve"
        assert np.array_equal(s._state["pos"].cpu().numpy(), g[k + "_pos"]), "pso positions"
        C = grid.shape[1]
        assert np.array_equal(np.array([r * C + c for r, c in res[0]], np.int32), g[k + "_best"])
        k = "ga_" + name
        N, K, seed = (int(x) for x in g[k + "_meta"])
        ga = GASolver(grid, num_generations=K, population_size=N, num_waypoints_per_chromosome=5, mutation_rate=0.1,
                      crossover_rate=0.8, tournament_size=3, rng_seed=seed, device=local, verbose=False,
                      group=dist.group.WORLD, **POLICY)
        res = ga.solve()
        assert np.array_equal(np.array(ga.convergence_curve), g[k + "_curve"]), "ga curve"
        assert np.array_equal(ga._pop["chrom"].cpu().numpy(), g[k + "_chrom"]), "ga population"
    from maaco_path_planing_b200.mpa import MPA
    for name in ("fig7_20", "blocks40_24"):
        k = "mpa_" + name
        N, K, seed, beta10 = (int(x) for x in g[k + "_meta"])
        grid = g[k + "_grid"].astype(int)
        m = MPA(grid, num_predators=N, num_iterations=K, FADs_rate=0.2, P_const=0.5, levy_beta=beta10 / 10.0,
                turn_penalty_factor=0.1, safety_penalty_factor=0.8, min_safe_distance=1.8, diagonal_obstacle_penalty=100.0,
                rng_seed=seed, device=local, verbose=False, group=dist.group.WORLD)
        m.solve_path_planning()
        curve = np.array([np.inf if v is None else v for v in m.convergence_curve_data])
        assert np.array_equal(curve, g[k + "_curve"]), "mpa curve"
        assert np.array_equal(m._pop["stats"][:, 4].cpu().numpy(), g[k + "_pop_fit"]), "mpa population"
    dist.barrier()
    if dist.get_rank() == 0:
        print(f"multigpu_solvers_check ok: world={dist.get_world_size()} PSO/GA/MPA sharded populations bit-exact vs reference goldens")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
