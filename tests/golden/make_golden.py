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


def main():
    grids = env_grids()
    np.savez_compressed(os.path.join(HERE, "env_grids.npz"), **grids)
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


if __name__ == "__main__":
    main()
