"""Times the UNMODIFIED reference's own Python hot path on one host core -- TEST / BENCH INFRASTRUCTURE ONLY
(bench.py's cpu_baseline leg starts one of these per host core; never imported by the product).

    python oracle/ref_timing.py maaco   <size> <map_seed> <seconds>
    python oracle/ref_timing.py fitness <size> <map_seed> <seconds>
    python oracle/ref_timing.py mpa     <size> <map_seed> <seconds>

The reference tree is found by ref_harness.ref_dir() ($MAACO_REF_DIR, /root/reference, baseline/_ref).  Native RNG
(the reference never seeds it): this is a timing of the stock code path, not a parity run.  Prints one JSON line.

  maaco    MAACO._construct_ant_solution_maaco (MAACO.py:278-302) over and over + ONE
           _update_pheromone_trails_maaco (MAACO.py:304-332) whose time is charged per tour
  fitness  GASolver._reconstruct_path_from_chromosome + _calculate_stats_for_path (ga_solver.py:58-93,
           helper.py:98-113) on random free-cell chromosomes, W = 5 (main.py:95-102)
  mpa      MPA._reconstruct_path_segment + the FADs step (MPA.py:284-318, :387-410) = one predator-iteration,
           driven through MPA.solve_path_planning's own loop on a small population
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as H  # noqa: E402

MAACO_PARAMS = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2,
                    k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)   # main.py:34-38
POLICY = dict(turn_penalty_factor=0.3, safety_penalty_factor=0.8, min_safe_distance=1.8,
              diagonal_obstacle_penalty_value=100.0)                                # main.py:21-24


def main():
    kind, size, seed, seconds = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    ref = H.load_reference()
    grid = H.blocks_map(size, 0.20, seed)
    out = {"kind": kind, "size": size}
    if kind == "maaco":
        with H.quiet():
            s = ref.MAACO.MAACO(np.array(grid), num_ants=4096, num_iterations=100, **MAACO_PARAMS)
        paths, steps = [], 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            p, ln, tn = s._construct_ant_solution_maaco(len(paths), 1)
            paths.append((p, ln, tn))
            steps += max(0, len(p) - 1)
        t_tours = time.perf_counter() - t0
        t1 = time.perf_counter()
        s._update_pheromone_trails_maaco(paths, min((x[1] for x in paths), default=float("inf")))
        t_upd = time.perf_counter() - t1
        out.update(evals=len(paths), seconds=t_tours + t_upd, tour_seconds=t_tours, update_seconds=t_upd,
                   successful=sum(1 for x in paths if x[0]), path_cells=sum(len(x[0]) for x in paths))
    elif kind == "fitness":
        with H.quiet():
            ga = ref.ga_solver.GASolver(grid=np.array(grid), num_generations=1, population_size=2,
                                        num_waypoints_per_chromosome=5, mutation_rate=0.1, crossover_rate=0.8,
                                        tournament_size=3, allow_diagonal_moves=True,
                                        restrict_diagonal_near_obstacle_policy=True, **POLICY)
        n, valid = 0, 0
        t0 = time.perf_counter()
        while n == 0 or time.perf_counter() - t0 < seconds:
            chrom = ga._create_chromosome()
            with H.quiet():
                path = ga._reconstruct_path_from_chromosome(chrom)
                ga._calculate_stats_for_path(path)
            n += 1
            valid += bool(path)
        out.update(evals=n, seconds=time.perf_counter() - t0, valid=valid)
    elif kind == "mpa":
        N = 8
        kw = dict(num_predators=N, num_iterations=1000000, FADs_rate=0.2, P_const=0.5, levy_beta=2.0,
                  turn_penalty_factor=0.1, safety_penalty_factor=0.8, min_safe_distance=1.8,
                  diagonal_obstacle_penalty=100.0, allow_diagonal_moves=True, restrict_diagonal_near_obstacle=True)
        with H.quiet():
            s = ref.MPA.MPA(grid=np.array(grid), **kw)
        # the reference's own loop (MPA.py:320-448), 12 iterations at a time (phases 1-3), until the clock runs out
        iters = 0
        s.num_iterations = 12
        t0 = time.perf_counter()
        with H.quiet():
            while iters == 0 or time.perf_counter() - t0 < seconds:
                s.convergence_curve_data = []
                s.solve_path_planning()
                iters += s.num_iterations
        out.update(evals=iters * N, seconds=time.perf_counter() - t0)
    else:
        raise SystemExit("kind must be maaco | fitness | mpa")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
