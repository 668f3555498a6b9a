"""Generate the committed golden vectors from the UNMODIFIED reference (run in the authoring
container only: needs /root/reference).  The reference ships no tests or fixtures, so these
vectors -- the reference's own outputs under the injected Philox streams of the RNG contract --
are what pins the C oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as H  # noqa: E402

MAACO_DEFAULT = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2,
                     k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)   # main.py:34-38


def env_grids():
    ref = H.load_reference()
    e = ref.env
    g7 = np.array(e.grid_fig7_layout_data)
    g7[0, 0] = 2
    g7[19, 19] = 3                                                                  # main.py:27-32
    out = {"fig7": g7, "fig13": np.array(e.grid_map_fig13_base_data),
           "image1": np.array(e.grid_map_from_image_data), "image2": np.array(e.grid_map_from_image_data2),
           "image3": np.array(e.grid_map_from_image_data3), "image5": np.array(e.grid_map_from_image_data5)}
    return {k: v.astype(np.uint8) for k, v in out.items()}


def env_raw_data():
    """The reference's demo maps exactly as env.py defines them (before main.py marks start/target), stored as
    data for the drop-in `env` module (maaco_path_planing_b200/dropin/env.py): key = the reference's variable name."""
    e = H.load_reference().env
    names = ["grid_fig7_layout_data", "grid_map_fig13_base_data", "grid_map_from_image_data",
             "grid_map_from_image_data2", "grid_map_from_image_data3", "grid_map_from_image_data5"]
    out = {n: np.array(getattr(e, n)).astype(np.uint8) for n in names}
    d = os.path.join(ROOT, "maaco_path_planing_b200", "data")
    os.makedirs(d, exist_ok=True)
    np.savez_compressed(os.path.join(d, "env_grids.npz"), **out)
    print("env data", {k: v.shape for k, v in out.items()})


def maaco_case(name, grid, N, K, seed, params):
    out = H.run_maaco(grid, dict(num_ants=N, num_iterations=K, **params), seed=seed)
    C = grid.shape[1]
    ncell = np.zeros((K, N), np.int32)
    length = np.zeros((K, N))
    turns = np.zeros((K, N), np.int32)
    flat = []
    for it in range(K):
        for a, (path, ln, tn) in enumerate(out["trace"]["tours"][it]):
            ncell[it, a] = len(path)
            length[it, a] = ln
            turns[it, a] = -1 if tn == float("inf") else tn
            flat.extend(r * C + c for r, c in path)
    res = out["result"]
    np.savez_compressed(
        os.path.join(HERE, f"maaco_{name}.npz"), grid=grid.astype(np.uint8), N=N, K=K, seed=seed,
        params=json.dumps(params), n_cells=ncell, length=length, turns=turns, cells=np.array(flat, np.int32),
        tau0=out["tau0"], tau=np.stack(out["trace"]["tau"]),
        best_cells=np.array([r * C + c for r, c in res[0]], np.int32), best_len=res[1],
        best_turns=-1 if res[2] == float("inf") else res[2],
        curve=np.array([np.inf if v is None else v for v in out["solver"].convergence_curve_data]))
    print(name, "tours", N * K, "succ", int((ncell > 0).sum()), "best", res[1], res[2])


def rand_grid(rng, n, m, dens):
    g = (rng.random((n, m)) < dens).astype(int)
    free = np.argwhere(g == 0)
    s = free[rng.integers(len(free))]
    g[s[0], s[1]] = 2
    free = np.argwhere(g == 0)
    t = free[rng.integers(len(free))]
    g[t[0], t[1]] = 3
    return g


def astar_cases(n_cases=240):
    """Reference outputs of AStarSolver.solve (variant 0) and MPA._a_star (variant 1) on random
    (grid, src, dst, avoid-set) tuples incl. unreachable targets, src==dst, obstacle endpoints."""
    ref = H.load_reference()
    rng = np.random.default_rng(77)
    rec = dict(grid=[], shape=[], src=[], dst=[], avoid=[], avoid_off=[0], flags=[], path0=[], off0=[0], path1=[],
               off1=[0], g1=[])
    for case in range(n_cases):
        n, m = int(rng.integers(5, 30)), int(rng.integers(5, 30))
        g = rand_grid(rng, n, m, rng.uniform(0.05, 0.35))
        allow_diag, restrict = bool(rng.random() < 0.9), bool(rng.random() < 0.8)
        cells_all = [(r, c) for r in range(n) for c in range(m)]

        def pick():
            if rng.random() < 0.08:
                return cells_all[rng.integers(len(cells_all))]
            fr = np.argwhere(g != 1)
            q = fr[rng.integers(len(fr))]
            return (int(q[0]), int(q[1]))
        src, dst = pick(), pick()
        if rng.random() < 0.04:
            dst = src
        avoid = {cells_all[rng.integers(len(cells_all))] for _ in range(int(rng.integers(0, n + m)))}
        if rng.random() < 0.1:
            avoid.add(dst)
        if rng.random() < 0.1:
            avoid.add(src)
        with H.quiet():
            solver = ref.astar.AStarSolver(grid=g, turn_penalty_factor=0, safety_penalty_factor=0, min_safe_distance=0,
                                           allow_diagonal_moves=allow_diag,
                                           restrict_diagonal_near_obstacle_policy=restrict,
                                           diagonal_obstacle_penalty_value=0)
            p0 = solver.solve(start_node_override=src, target_node_override=dst, nodes_to_avoid=set(avoid))[0]
        mpa = ref.MPA.MPA.__new__(ref.MPA.MPA)
        mpa.grid = np.array(g, dtype=int)
        mpa.rows, mpa.cols = n, m
        mpa.allow_diagonal_moves, mpa.restrict_diagonal_near_obstacle = allow_diag, restrict
        p1, g1 = mpa._a_star(src, dst, set(avoid))
        rec["grid"].append(g.astype(np.uint8).ravel())
        rec["shape"].append((n, m))
        rec["src"].append(src[0] * m + src[1])
        rec["dst"].append(dst[0] * m + dst[1])
        av = sorted(r * m + c for r, c in avoid)
        rec["avoid"].extend(av)
        rec["avoid_off"].append(len(rec["avoid"]))
        rec["flags"].append((int(allow_diag), int(restrict)))
        rec["path0"].extend(int(r) * m + int(c) for r, c in p0)
        rec["off0"].append(len(rec["path0"]))
        rec["path1"].extend(int(r) * m + int(c) for r, c in p1)
        rec["off1"].append(len(rec["path1"]))
        rec["g1"].append(float(g1))
    np.savez_compressed(os.path.join(HERE, "astar_cases.npz"), grid=np.concatenate(rec["grid"]),
                        shape=np.array(rec["shape"], np.int32), src=np.array(rec["src"], np.int32),
                        dst=np.array(rec["dst"], np.int32), avoid=np.array(rec["avoid"], np.int32),
                        avoid_off=np.array(rec["avoid_off"], np.int64), flags=np.array(rec["flags"], np.int32),
                        path0=np.array(rec["path0"], np.int32), off0=np.array(rec["off0"], np.int64),
                        path1=np.array(rec["path1"], np.int32), off1=np.array(rec["off1"], np.int64),
                        g1=np.array(rec["g1"]))
    print("astar cases", n_cases, "variant0 fails", sum(1 for i in range(n_cases) if rec["off0"][i] == rec["off0"][i + 1]))


POLICY = dict(turn_penalty_factor=0.3, safety_penalty_factor=0.8, min_safe_distance=1.8,
              diagonal_obstacle_penalty_value=100.0)                                # main.py:21-24


def fitness_cases():
    """Reference GA-style waypoint chains + stats (ga_solver.py:58-93, helper.py:98-113) and MPA-mode stats."""
    ref = H.load_reference()
    rng = np.random.default_rng(78)
    grids = {"fig7": env_grids()["fig7"].astype(int), "blocks40": H.blocks_map(40, 0.2, 21),
             "blocks64": H.blocks_map(64, 0.2, 22), "rect": H.blocks_map(0, 0.18, 23, rows=30, cols=48)}
    out = {}
    for name, g in grids.items():
        R, Cc = g.shape
        for msd in (1.8, 2.5):
            pol = dict(POLICY, min_safe_distance=msd)
            with H.quiet():
                ga = ref.ga_solver.GASolver(grid=g, num_generations=1, population_size=2, num_waypoints_per_chromosome=5,
                                            mutation_rate=0.1, crossover_rate=0.8, tournament_size=3,
                                            allow_diagonal_moves=True, restrict_diagonal_near_obstacle_policy=True, **pol)
            free = np.argwhere(g != 1)
            N, W = 12, 5
            wps = np.zeros((N, W), np.int32)
            paths, offs, stats, mstats = [], [0], [], []
            for i in range(N):
                wp = [tuple(int(x) for x in free[rng.integers(len(free))]) for _ in range(W)]
                if i % 5 == 4:
                    wp[2] = (int(rng.integers(R)), int(rng.integers(Cc)))              # may be an obstacle (PSO style)
                if i % 6 == 5:
                    wp[1] = wp[0]                                                    # duplicate waypoint
                wps[i] = [r * Cc + c for r, c in wp]
                with H.quiet():
                    path = ga._reconstruct_path_from_chromosome(list(wp))
                    st = ga._calculate_stats_for_path(path)
                paths.extend(int(r) * Cc + int(c) for r, c in path)
                offs.append(len(paths))
                stats.append([float(st[1]), float(st[2]), float(st[3]), float(st[4]), float(st[5])])
                mpa = ref.MPA.MPA.__new__(ref.MPA.MPA)
                mpa.grid = np.array(g, dtype=int)
                mpa.rows, mpa.cols = R, Cc
                mpa.restrict_diagonal_near_obstacle = True
                mpa.diagonal_obstacle_penalty_val = 100.0
                mpa.turn_penalty_factor_mpa, mpa.safety_penalty_factor_mpa = 0.1, 0.8
                ms = mpa._calculate_path_stats(path)
                mstats.append([float(x) for x in ms[1:]])
            key = f"{name}_msd{msd}"
            out[key + "_grid"] = g.astype(np.uint8)
            out[key + "_wps"] = wps
            out[key + "_paths"] = np.array(paths, np.int32)
            out[key + "_offs"] = np.array(offs, np.int64)
            out[key + "_stats"] = np.array(stats)
            out[key + "_mpastats"] = np.array(mstats)
            print(key, "valid", sum(1 for i in range(N) if offs[i + 1] > offs[i]), "/", N)
    np.savez_compressed(os.path.join(HERE, "fitness_cases.npz"), **out)


def solver_cases():
    """Whole-solver trajectories of the reference PSO and GA under the tape (main.py:93-118 parameters,
    reduced population / iteration counts so the pure-Python reference finishes in seconds)."""
    grids = {"fig7": env_grids()["fig7"].astype(int), "blocks40": H.blocks_map(40, 0.2, 31)}
    out = {}
    for name, g in grids.items():
        for N, K, seed in ((20, 6, 301), (33, 4, 302)):
            kw = dict(num_iterations=K, num_particles=N, num_waypoints_per_particle=5, w=0.7, c1=1.5, c2=1.5,
                      allow_diagonal_moves=True, restrict_diagonal_near_obstacle_policy=True, **POLICY)
            r = H.run_pso(g, kw, seed)
            k = f"pso_{name}_{N}"
            out[k + "_grid"] = g.astype(np.uint8)
            out[k + "_meta"] = np.array([N, K, seed])
            out[k + "_curve"] = np.array(r["curve"])
            out[k + "_stats"] = np.array([float(x) for x in r["result"][1:]])
            out[k + "_best"] = r["best_cells"]
            for f in ("pos", "vel", "pbest_fit", "cur_fit"):
                out[k + "_" + f] = r[f]
            out[k + "_init_pos"] = r["init"]["pos"]
            out[k + "_init_vel"] = r["init"]["vel"]
            out[k + "_init_fit"] = r["init"]["fit"]
            print(k, "curve", r["curve"][0], "->", r["curve"][-1], "draws", r["draws"])
            kw = dict(num_generations=K, population_size=N, num_waypoints_per_chromosome=5, mutation_rate=0.1,
                      crossover_rate=0.8, tournament_size=3, allow_diagonal_moves=True,
                      restrict_diagonal_near_obstacle_policy=True, **POLICY)
            r = H.run_ga(g, kw, seed + 50)
            k = f"ga_{name}_{N}"
            out[k + "_grid"] = g.astype(np.uint8)
            out[k + "_meta"] = np.array([N, K, seed + 50])
            out[k + "_curve"] = np.array(r["curve"])
            out[k + "_stats"] = np.array([float(x) for x in r["result"][1:]])
            out[k + "_best"] = r["best_cells"]
            out[k + "_chrom"] = r["chrom"]
            out[k + "_fit"] = r["fit"]
            out[k + "_init_chrom"] = r["init"]["chrom"]
            out[k + "_init_fit"] = r["init"]["fit"]
            print(k, "curve", r["curve"][0], "->", r["curve"][-1], "draws", r["draws"])
    # MPA (main.py:44-52 parameters; levy_beta=2.0 is the reference's degenerate Levy step, 1.5 a real one)
    for name, g in grids.items():
        for N, K, beta, seed in ((20, 12, 2.0, 401), (24, 9, 1.5, 402)):
            kw = dict(num_predators=N, num_iterations=K, FADs_rate=0.2, P_const=0.5, levy_beta=beta,
                      turn_penalty_factor=0.1, safety_penalty_factor=0.8, min_safe_distance=1.8,
                      diagonal_obstacle_penalty=100.0, allow_diagonal_moves=True, restrict_diagonal_near_obstacle=True)
            r = H.run_mpa(g, kw, seed)
            k = f"mpa_{name}_{N}"
            out[k + "_grid"] = g.astype(np.uint8)
            out[k + "_meta"] = np.array([N, K, seed, int(beta * 10)])
            out[k + "_curve"] = np.array(r["curve"])
            out[k + "_stats"] = np.array([float(x) for x in r["result"][1:]])
            out[k + "_best"] = r["best_cells"]
            out[k + "_pop_cells"] = r["pop_cells"]
            out[k + "_pop_offs"] = r["pop_offs"]
            out[k + "_pop_fit"] = r["pop_fit"]
            print(k, "curve", r["curve"][0], "->", r["curve"][-1], "draws", r["draws"])
    np.savez_compressed(os.path.join(HERE, "solver_cases.npz"), **out)


def snake_map(n=9):
    """One-cell-wide corridor from (0,0) to (n-1,n-1): almost no random waypoint chain is feasible (a segment may not
    re-enter the cells of the earlier ones), the direct start -> target search is."""
    g = np.ones((n, n), int)
    for r in range(0, n, 2):
        g[r, :] = 0
        if r + 1 < n:
            g[r + 1, n - 1 if (r // 2) % 2 == 0 else 0] = 0
    g[0, 0], g[n - 1, n - 1] = 2, 3
    return g


def fallback_cases():
    """PSO / GA runs of the reference whose initialisation finds no valid waypoint chain in N*20 attempts and
    falls back to the direct-path individual (pso.py:128-145, ga_solver.py:111-117); plus the total failure
    (start walled in: pso.py:147-157, ga_solver.py:120-126)."""
    g = snake_map(9)
    out = {"grid": g.astype(np.uint8)}
    seed = 500
    while True:
        kw = dict(num_iterations=3, num_particles=3, num_waypoints_per_particle=5, w=0.7, c1=1.5, c2=1.5,
                  allow_diagonal_moves=True, restrict_diagonal_near_obstacle_policy=True, **POLICY)
        r = H.run_pso(g, kw, seed)
        if np.all(r["init"]["pos"] == 0.0) and r["result"][0]:
            break
        seed += 1
    out["pso_meta"] = np.array([3, 3, seed])
    out["pso_curve"] = np.array(r["curve"])
    out["pso_stats"] = np.array([float(x) for x in r["result"][1:]])
    out["pso_best"] = r["best_cells"]
    out["pso_pos"] = r["pos"]
    out["pso_cur_fit"] = r["cur_fit"]
    print("pso fallback seed", seed, r["result"][1:], r["curve"])
    seed = 600
    while True:
        kw = dict(num_generations=3, population_size=2, num_waypoints_per_chromosome=5, mutation_rate=0.1,
                  crossover_rate=0.8, tournament_size=3, allow_diagonal_moves=True,
                  restrict_diagonal_near_obstacle_policy=True, **POLICY)
        r = H.run_ga(g, kw, seed)
        if r["init"]["chrom"].size == 0 and r["result"][0]:
            break
        seed += 1
    out["ga_meta"] = np.array([2, 3, seed])
    out["ga_curve"] = np.array(r["curve"])
    out["ga_stats"] = np.array([float(x) for x in r["result"][1:]])
    out["ga_best"] = r["best_cells"]
    out["ga_fit"] = r["fit"]
    print("ga fallback seed", seed, r["result"][1:], r["curve"])
    # total failure: the start is walled in
    w = snake_map(9)
    w[0, 1] = 1
    w[1, 0] = 1
    w[1, 1] = 1
    out["walled_grid"] = w.astype(np.uint8)
    kw = dict(num_iterations=2, num_particles=2, num_waypoints_per_particle=3, w=0.7, c1=1.5, c2=1.5, **POLICY)
    r = H.run_pso(w, kw, 701)
    out["walled_pso_stats"] = np.array([float(x) for x in r["result"][1:]])
    kw = dict(num_generations=2, population_size=2, num_waypoints_per_chromosome=3, mutation_rate=0.1,
              crossover_rate=0.8, **POLICY)
    r2 = H.run_ga(w, kw, 702)
    out["walled_ga_stats"] = np.array([float(x) for x in r2["result"][1:]])
    print("walled", r["result"], r2["result"])
    np.savez_compressed(os.path.join(HERE, "fallback_cases.npz"), **out)


def dijkstra_cases(n_cases=120):
    """Reference outputs of DijkstraSolver.solve (dijkstra.py:32-97) on random (grid, src, dst, avoid) tuples."""
    ref = H.load_reference()
    rng = np.random.default_rng(5)
    rec = dict(grid=[], shape=[], src=[], dst=[], avoid=[], avoid_off=[0], flags=[], path=[], off=[0])
    for case in range(n_cases):
        n, m = int(rng.integers(5, 30)), int(rng.integers(5, 30))
        g = rand_grid(rng, n, m, rng.uniform(0.05, 0.35))
        ad, rs = bool(rng.random() < 0.9), bool(rng.random() < 0.8)
        cells = [(r, c) for r in range(n) for c in range(m)]

        def pick():
            if rng.random() < 0.08:
                return cells[rng.integers(len(cells))]
            fr = np.argwhere(g != 1)
            q = fr[rng.integers(len(fr))]
            return (int(q[0]), int(q[1]))
        src, dst = pick(), pick()
        avoid = {cells[rng.integers(len(cells))] for _ in range(int(rng.integers(0, n + m)))}
        with H.quiet():
            sol = ref.dijkstra.DijkstraSolver(grid=g, turn_penalty_factor=0, safety_penalty_factor=0, min_safe_distance=0,
                                              allow_diagonal_moves=ad, restrict_diagonal_near_obstacle_policy=rs,
                                              diagonal_obstacle_penalty_value=0)
            rp = sol.solve(start_node_override=src, target_node_override=dst, nodes_to_avoid=set(avoid))[0]
        rec["grid"].append(g.astype(np.uint8).ravel())
        rec["shape"].append((n, m))
        rec["src"].append(src[0] * m + src[1])
        rec["dst"].append(dst[0] * m + dst[1])
        rec["avoid"].extend(sorted(r * m + c for r, c in avoid))
        rec["avoid_off"].append(len(rec["avoid"]))
        rec["flags"].append((int(ad), int(rs)))
        rec["path"].extend(int(r) * m + int(c) for r, c in rp)
        rec["off"].append(len(rec["path"]))
    np.savez_compressed(os.path.join(HERE, "dijkstra_cases.npz"), grid=np.concatenate(rec["grid"]),
                        shape=np.array(rec["shape"], np.int32), src=np.array(rec["src"], np.int32),
                        dst=np.array(rec["dst"], np.int32), avoid=np.array(rec["avoid"], np.int32),
                        avoid_off=np.array(rec["avoid_off"], np.int64), flags=np.array(rec["flags"], np.int32),
                        path=np.array(rec["path"], np.int32), off=np.array(rec["off"], np.int64))
    print("dijkstra cases", n_cases)


def maaco_extra_cases():
    """Orientations other than start-top-left / target-bottom-right: a start and target on one row (the
    orientation filter MAACO.py:146-157 then leaves five moves, not three) and a target up-left of the start
    behind walls with gaps (other quadrant; dead ends force strategies 2 and 3, :169-180)."""
    row = np.zeros((14, 30), int)
    row[3:11, 12] = 1
    row[0:6, 20] = 1
    row[8:14, 24] = 1
    row[7, 0], row[7, 29] = 2, 3
    maaco_case("aligned_row", row, 24, 5, 107, MAACO_DEFAULT)
    rev = H.blocks_map(0, 0.18, 21, rows=40, cols=70)
    rev[rev == 2] = 0
    rev[rev == 3] = 0
    rev[20, 10:60] = 1
    rev[20, 33:36] = 0
    rev[39, 69], rev[0, 0] = 2, 3
    maaco_case("reverse_diag", rev, 32, 4, 108, MAACO_DEFAULT)


def maaco_default_cases():
    """BASELINE config 1 and its siblings: MAACO at main.py's own parameters (main.py:34-38: 50 ants x 100
    iterations) on every 20x20 demo map main.py / env.py define, and 32 ants x 3 iterations on the 256x256
    grid_map_from_image_data5 (env.py:114-371)."""
    grids = env_grids()
    for name in ("fig7", "fig13", "image1", "image2", "image3"):
        maaco_case(name + "_default", grids[name].astype(int), 50, 100, 120 + len(name), MAACO_DEFAULT)
    maaco_case("image5", grids["image5"].astype(int), 32, 3, 131, MAACO_DEFAULT)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "maaco_default":
        maaco_default_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "dijkstra":
        dijkstra_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "fallback":
        fallback_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "env_raw":
        env_raw_data()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "maaco_extra":
        maaco_extra_cases()
        return
    dijkstra_cases()
    fallback_cases()
    solver_cases()
    astar_cases()
    fitness_cases()
    grids = env_grids()
    np.savez_compressed(os.path.join(HERE, "env_grids.npz"), **grids)
    env_raw_data()
    maaco_case("fig7", grids["fig7"].astype(int), 20, 8, 101, MAACO_DEFAULT)
    maaco_case("fig13", grids["fig13"].astype(int), 16, 5, 102, MAACO_DEFAULT)
    maaco_case("blocks64", H.blocks_map(64, 0.2, 7), 32, 4, 103, MAACO_DEFAULT)
    near = np.zeros((12, 12), int)
    near[5, 5] = 2
    near[7, 8] = 3
    near[6, 6] = 1
    maaco_case("near_roulette", near, 24, 6, 104, dict(MAACO_DEFAULT, beta=2.0))
    maaco_case("near_alpha2", near, 24, 4, 105, dict(MAACO_DEFAULT, beta=1.5, alpha=2.0, Q=50.0))
    rect = H.blocks_map(0, 0.15, 9, rows=24, cols=40)
    maaco_case("rect24x40", rect, 16, 4, 106, MAACO_DEFAULT)
    maaco_extra_cases()
    maaco_default_cases()


if __name__ == "__main__":
    main()
