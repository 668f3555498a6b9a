"""main.py-equivalent driver: runs MAACO, MPA, A*, Dijkstra, GA and PSO on one map with
the parameter blocks of the reference's demo (main.py:21-24, 34-52, 93-118) and exports paths, statistics
and convergence data to .npz / prints a table -- no display or matplotlib needed.

    python -m maaco_path_planing_b200.demo --map blocks:64 --seed 1 --out results.npz
    python -m maaco_path_planing_b200.demo --map npz:tests/golden/env_grids.npz:fig7
"""
from __future__ import annotations

import argparse
import time

import numpy as np

COMMON = dict(turn_penalty_factor=0.3, safety_penalty_factor=0.8, min_safe_distance=1.8)     # main.py:21-23
DIAG = 100.0                                                                                 # main.py:24


def load_map(spec):
    from .gridmap import blocks_map
    kind, _, rest = spec.partition(":")
    if kind == "blocks":
        n, _, seed = rest.partition(":")
        return blocks_map(int(n), 0.20, seed=int(seed or 0))
    if kind == "npz":
        path, _, key = rest.rpartition(":")
        return np.load(path)[key].astype(int)
    raise ValueError("--map blocks:<n>[:seed] | npz:<file>:<key>")


def run(grid, seed=0, scale=1.0, verbose=False):
    from . import MAACO
    from .astar import AStarSolver
    from .dijkstra import DijkstraSolver
    from .ga_solver import GASolver
    from .mpa import MPA
    from .pso import PSOSolver
    it = lambda n: max(2, int(n * scale))
    out = {}

    def timed(name, fn):
        t0 = time.perf_counter()
        res = fn()
        out[name] = {"result": res, "seconds": time.perf_counter() - t0}
        return res

    m = MAACO(grid, num_ants=50, num_iterations=it(100), alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0,
              wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1,
              rng_seed=seed, verbose=verbose)                                                # main.py:34-38
    timed("MAACO", m.solve_path_planning)
    out["MAACO"]["curve"] = m.convergence_curve_data
    p = MPA(grid, num_predators=50, num_iterations=it(100), FADs_rate=0.2, P_const=0.5, levy_beta=2.0,
            turn_penalty_factor=0.1, safety_penalty_factor=COMMON["safety_penalty_factor"],
            min_safe_distance=COMMON["min_safe_distance"], diagonal_obstacle_penalty=DIAG, allow_diagonal_moves=True,
            restrict_diagonal_near_obstacle=True, rng_seed=seed + 1, verbose=verbose)        # main.py:44-52
    timed("MPA", p.solve_path_planning)
    out["MPA"]["curve"] = p.convergence_curve_data
    pol = dict(COMMON, allow_diagonal_moves=True, restrict_diagonal_near_obstacle_policy=True,
               diagonal_obstacle_penalty_value=DIAG)
    timed("A*", AStarSolver(grid, **pol).solve)                                              # main.py:65-75
    timed("Dijkstra", DijkstraSolver(grid, **pol).solve)                                     # main.py:79-89
    g = GASolver(grid, num_generations=it(100), population_size=50, num_waypoints_per_chromosome=5, mutation_rate=0.1,
                 crossover_rate=0.8, tournament_size=3, rng_seed=seed + 2, verbose=verbose, **pol)   # main.py:93-103
    timed("GA", g.solve)
    out["GA"]["curve"] = g.convergence_curve
    s = PSOSolver(grid, num_iterations=it(50), num_particles=100, num_waypoints_per_particle=5, w=0.7, c1=1.5, c2=1.5,
                  rng_seed=seed + 3, verbose=verbose, **pol)                                 # main.py:109-118
    timed("PSO", s.solve)
    out["PSO"]["curve"] = s.convergence_curve
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default="blocks:40:1")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the demo's iteration counts")
    ap.add_argument("--out", default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    grid = load_map(a.map)
    res = run(grid, a.seed, a.scale, a.verbose)
    print(f"{'algorithm':10s} {'cells':>6s} {'length':>10s} {'turns':>6s} {'fitness':>10s} {'seconds':>8s}")
    save = {"grid": grid}
    for name, r in res.items():
        t = r["result"]
        fit = t[5] if len(t) > 3 else float("nan")
        print(f"{name:10s} {len(t[0]):6d} {t[1]:10.3f} {str(t[2]):>6s} {fit:10.3f} {r['seconds']:8.3f}")
        save[name + "_path"] = np.array(t[0], np.int32).reshape(-1, 2)
        save[name + "_stats"] = np.array([float(x) for x in t[1:]])
        if "curve" in r:
            save[name + "_curve"] = np.array([np.nan if v is None else v for v in r["curve"]])
    if a.out:
        np.savez_compressed(a.out, **save)
        print("wrote", a.out)


if __name__ == "__main__":
    main()
