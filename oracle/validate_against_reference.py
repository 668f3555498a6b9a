"""Bulk validation of the C oracle against the UNMODIFIED reference (authoring container only).
Writes oracle/VALIDATION.json.  Covers the connector A* (astar.py), MPA's private A* (MPA.py),
path statistics (helper.py / MPA.py) and the waypoint chain (pso.py / ga_solver.py).

    python oracle/validate_against_reference.py [n_cases]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pyoracle as O  # noqa: E402
import ref_harness as H  # noqa: E402


def rand_grid(rng, n, dens):
    g = (rng.random((n, n)) < dens).astype(int)
    free = np.argwhere(g == 0)
    s = free[rng.integers(len(free))]
    g[s[0], s[1]] = 2
    free = np.argwhere(g == 0)
    t = free[rng.integers(len(free))]
    g[t[0], t[1]] = 3
    return g


def main(n_cases=1500):
    ref = H.load_reference()
    rng = np.random.default_rng(2024)
    out = {}
    # ---------------- A1 / A2 ----------------
    bad1 = bad2 = 0
    fails1 = fails2 = 0
    t0 = time.time()
    for case in range(n_cases):
        n = int(rng.integers(6, 34))
        g = rand_grid(rng, n, rng.uniform(0.05, 0.35))
        C = n
        allow_diag = bool(rng.random() < 0.9)
        restrict = bool(rng.random() < 0.8)
        cells_all = [(r, c) for r in range(n) for c in range(n)]
        # endpoints: mostly free cells, sometimes obstacle / equal
        def pick():
            if rng.random() < 0.08:
                return cells_all[rng.integers(len(cells_all))]
            fr = np.argwhere(g != 1)
            p = fr[rng.integers(len(fr))]
            return (int(p[0]), int(p[1]))
        src, dst = pick(), pick()
        if rng.random() < 0.03:
            dst = src
        k = int(rng.integers(0, n * 2))
        avoid = set()
        for _ in range(k):
            avoid.add(cells_all[rng.integers(len(cells_all))])
        if rng.random() < 0.1:
            avoid.add(dst)
        if rng.random() < 0.1:
            avoid.add(src)
        bits = O.cells_to_bits([r * C + c for r, c in avoid], n * n)
        orc = O.AStarOracle(g, allow_diag, restrict)
        # A1
        with H.quiet():
            solver = ref.astar.AStarSolver(grid=g, turn_penalty_factor=0, safety_penalty_factor=0, min_safe_distance=0,
                                           allow_diagonal_moves=allow_diag, restrict_diagonal_near_obstacle_policy=restrict,
                                           diagonal_obstacle_penalty_value=0)
            rp = solver.solve(start_node_override=src, target_node_override=dst, nodes_to_avoid=set(avoid))[0]
        op, og, _, _ = orc.solve(0, src[0] * C + src[1], dst[0] * C + dst[1], bits)
        rpc = [int(r) * C + int(c) for r, c in rp]
        if rpc != list(op):
            bad1 += 1
        fails1 += (len(rpc) == 0)
        # A2
        mpa = ref.MPA.MPA.__new__(ref.MPA.MPA)
        mpa.grid = np.array(g, dtype=int)
        mpa.rows, mpa.cols = n, n
        mpa.allow_diagonal_moves = allow_diag
        mpa.restrict_diagonal_near_obstacle = restrict
        rp2, rg2 = mpa._a_star(src, dst, set(avoid))
        op2, og2, _, _ = orc.solve(1, src[0] * C + src[1], dst[0] * C + dst[1], bits)
        rpc2 = [int(r) * C + int(c) for r, c in rp2]
        if rpc2 != list(op2) or not (float(rg2) == og2):
            bad2 += 1
        fails2 += (len(rpc2) == 0)
    out["astar_py_variant0"] = {"cases": n_cases, "mismatches": bad1, "unreachable_or_invalid": fails1}
    out["mpa_astar_variant1"] = {"cases": n_cases, "mismatches": bad2, "unreachable_or_invalid": fails2}
    print("A*", out, f"{time.time()-t0:.1f}s", flush=True)

    # ---------------- waypoint chains + stats (GA / PSO style) ----------------
    bad_path = bad_stats = 0
    n_chain = max(60, n_cases // 10)
    invalid = 0
    for case in range(n_chain):
        n = int(rng.integers(10, 40))
        if rng.random() < 0.5:
            g = H.blocks_map(n, 0.2, int(rng.integers(1 << 30)))
        else:
            g = rand_grid(rng, n, rng.uniform(0.05, 0.25))
        W = int(rng.integers(1, 6))
        tpf, spf, msd, dv = 0.3, 0.8, float(rng.choice([1.8, 1.5, 2.5, 1.0, 3.2])), 100.0
        with H.quiet():
            ga = ref.ga_solver.GASolver(grid=g, num_generations=1, population_size=2, num_waypoints_per_chromosome=W,
                                        mutation_rate=0.1, crossover_rate=0.8, tournament_size=3,
                                        turn_penalty_factor=tpf, safety_penalty_factor=spf, min_safe_distance=msd,
                                        allow_diagonal_moves=True, restrict_diagonal_near_obstacle_policy=True,
                                        diagonal_obstacle_penalty_value=dv)
        free = np.argwhere(np.array(g) != 1)
        wp = [tuple(int(x) for x in free[rng.integers(len(free))]) for _ in range(W)]
        if rng.random() < 0.15:  # PSO-style: waypoint may be an obstacle / duplicate
            wp[0] = (int(rng.integers(n)), int(rng.integers(n)))
        with H.quiet():
            rpath = ga._reconstruct_path_from_chromosome(list(wp))
            rstats = ga._calculate_stats_for_path(rpath)
        C = n
        cells, ncell, stats, _ = O.waypoint_fitness(g, np.array([[r * C + c for r, c in wp]], np.int32), tpf, spf, msd, dv)
        rpc = [int(r) * C + int(c) for r, c in rpath]
        if rpc != list(cells[0, :ncell[0]]):
            bad_path += 1
            continue
        invalid += (len(rpc) == 0)
        want = [float(rstats[1]), float(rstats[2]), float(rstats[3]), float(rstats[4]), float(rstats[5])]
        if not all((a == b) for a, b in zip(want, stats[0].tolist())):
            bad_stats += 1
        # MPA-mode stats on the same path
        if rpc:
            mpa = ref.MPA.MPA.__new__(ref.MPA.MPA)
            mpa.grid = np.array(g, dtype=int)
            mpa.rows, mpa.cols = n, n
            mpa.restrict_diagonal_near_obstacle = True
            mpa.diagonal_obstacle_penalty_val = dv
            mpa.turn_penalty_factor_mpa, mpa.safety_penalty_factor_mpa = 0.1, spf
            ms = mpa._calculate_path_stats(rpath)
            os_ = O.path_stats(g, rpc, 0.1, spf, msd, dv, True, mode=1)
            if [float(x) for x in ms[1:]] != os_.tolist():
                bad_stats += 1
    out["waypoint_chain"] = {"cases": n_chain, "path_mismatches": bad_path, "stats_mismatches": bad_stats, "invalid_paths": invalid}
    print(out["waypoint_chain"], flush=True)
    if "--no-write" not in sys.argv:
        json.dump(out, open(os.path.join(HERE, "VALIDATION.json"), "w"), indent=1)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1500)
