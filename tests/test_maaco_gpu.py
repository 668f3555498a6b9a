"""Parity of the CUDA colony pass (through the C ABI) with the oracle and the reference's golden
vectors.  Cell sequences / chosen ants bit-exact; lengths and pheromone bit-exact fp64 (tolerance
stated where alpha != 1 uses the device pow)."""
import numpy as np
import pytest

from conftest import MAACO_CASES, MAACO_DEFAULT, MAACO_DEFAULT_CASES, load_golden

pytestmark = pytest.mark.gpu


def _check_pass(dev, orc, it, N, exact_tau=True, rtol=0.0):
    dev.run_iteration(it)
    nc, ln, tn, cells = dev.last_tours()
    ocells, onc, oln, otn, _ = orc.iterate(it)
    assert np.array_equal(nc, onc)
    assert np.array_equal(tn, otn)
    assert np.array_equal(ln, oln)
    for a in range(N):
        assert np.array_equal(cells[a, :nc[a]], ocells[a, :onc[a]]), f"ant {a} path differs (it {it})"
    tau = dev.pheromone_matrix.ravel()
    if exact_tau:
        assert np.array_equal(tau, orc.tau)
    else:
        np.testing.assert_allclose(tau, orc.tau, rtol=rtol, atol=0)


@pytest.mark.parametrize("name", MAACO_CASES)
@pytest.mark.parametrize("apw", [0, 1, 4, 32])
def test_golden_trajectories(name, apw):
    """CUDA vs the reference's own recorded trajectory (per-ant paths, tau after every pass)."""
    from maaco_path_planing_b200 import MAACO
    g = load_golden("maaco_" + name)
    N, K = int(g["N"]), int(g["K"])
    alpha1 = g["params"]["alpha"] == 1.0
    dev = MAACO(g["grid"].astype(int), N, K, rng_seed=int(g["seed"]), ants_per_warp=apw, verbose=False, **g["params"])
    assert np.array_equal(dev.pheromone_matrix, g["tau0"])
    pos = 0
    for it in range(1, K + 1):
        dev.run_iteration(it)
        nc, ln, tn, cells = dev.last_tours()
        if alpha1:
            assert np.array_equal(nc, g["n_cells"][it - 1])
            assert np.array_equal(ln, g["length"][it - 1])
            assert np.array_equal(tn, g["turns"][it - 1])
            for a in range(N):
                assert np.array_equal(cells[a, :nc[a]], g["cells"][pos:pos + nc[a]])
                pos += nc[a]
            assert np.array_equal(dev.pheromone_matrix, g["tau"][it - 1])
        else:
            # alpha != 1: tau**alpha is CUDA's pow (<= 2 ulp from glibc's, include/mpp.h "alpha != 1"), which can only
            # flip a choice between two moves whose attractiveness agrees to ~1e-16 relative.  On this fixture every
            # pass agrees on every discrete result; tau then differs only through ... nothing (deposits are Q / length),
            # so it is compared exactly as well.
            assert np.array_equal(nc, g["n_cells"][it - 1])
            assert np.array_equal(tn, g["turns"][it - 1])
            assert np.array_equal(ln, g["length"][it - 1])
            for a in range(N):
                assert np.array_equal(cells[a, :nc[a]], g["cells"][pos:pos + nc[a]])
                pos += nc[a]
            assert np.array_equal(dev.pheromone_matrix, g["tau"][it - 1])


@pytest.mark.parametrize("name", MAACO_DEFAULT_CASES)
@pytest.mark.parametrize("apw", [0, 32])
def test_reference_default_trajectories(name, apw):
    """BASELINE config 1 (MAACO at main.py:34-38's parameters: 50 ants x 100 iterations) on each demo map of env.py,
    and grid_map_from_image_data5 (256x256): every tour of every pass, tau after every pass, the returned best path and
    the convergence curve equal the reference's own recorded run."""
    from maaco_path_planing_b200 import MAACO
    g = load_golden("maaco_" + name)
    N, K = int(g["N"]), int(g["K"])
    dev = MAACO(g["grid"].astype(int), N, K, rng_seed=int(g["seed"]), ants_per_warp=apw, verbose=False, **g["params"])
    assert np.array_equal(dev.pheromone_matrix, g["tau0"])
    pos = 0
    for it in range(1, K + 1):
        dev.run_iteration(it)
        nc, ln, tn, cells = dev.last_tours()
        assert np.array_equal(nc, g["n_cells"][it - 1])
        assert np.array_equal(ln, g["length"][it - 1])
        assert np.array_equal(tn, g["turns"][it - 1])
        for a in range(N):
            assert np.array_equal(cells[a, :nc[a]], g["cells"][pos:pos + nc[a]])
            pos += nc[a]
        assert np.array_equal(dev.pheromone_matrix, g["tau"][it - 1])
    path, length, turns = dev.solve_path_planning()
    C = g["grid"].shape[1]
    assert [r * C + c for r, c in path] == g["best_cells"].tolist()
    assert (length == float(g["best_len"])) and (turns == int(g["best_turns"]) if int(g["best_turns"]) >= 0 else turns == float("inf"))
    curve = np.array([np.inf if v is None else v for v in dev.convergence_curve_data])
    assert np.array_equal(curve, g["curve"])


@pytest.mark.parametrize("apw", [0, 1, 8, 32])
def test_solve_matches_oracle_and_reference_api(apw):
    from maaco_path_planing_b200 import MAACO
    import pyoracle as O
    g = load_golden("maaco_fig7")
    N, K, seed = 50, 30, 77
    dev = MAACO(g["grid"].astype(int), N, K, rng_seed=seed, ants_per_warp=apw, verbose=False, **MAACO_DEFAULT)
    path, length, turns = dev.solve_path_planning()
    orc = O.MaacoOracle(g["grid"].astype(int), N, K, seed=seed, **MAACO_DEFAULT)
    opath, olen, oturns = orc.solve()
    assert [r * 20 + c for r, c in path] == list(opath)
    assert length == olen and turns == oturns
    assert dev.convergence_curve_data == orc.curve
    assert np.array_equal(dev.pheromone_matrix.ravel(), orc.tau)
    assert dev.best_path_overall[0] == dev.start_node and dev.best_path_overall[-1] == dev.target_node


@pytest.mark.parametrize("size,N,seed", [(100, 256, 1), (256, 512, 2)])
def test_block_maps_vs_oracle(size, N, seed):
    from maaco_path_planing_b200 import MAACO, blocks_map
    import pyoracle as O
    g = blocks_map(size, 0.2, seed=1000 + seed)
    dev = MAACO(g, N, 3, rng_seed=seed, verbose=False, **MAACO_DEFAULT)
    orc = O.MaacoOracle(g, N, 3, seed=seed, threads=0, **MAACO_DEFAULT)
    for it in (1, 2, 3):
        _check_pass(dev, orc, it, N)


def test_config4_full_size_vs_oracle_and_properties():
    """BASELINE config 4: 4096 ants on 512x512.  Oracle parity on the full colony (the C oracle
    finishes in well under a second per pass) plus size-independent properties."""
    from maaco_path_planing_b200 import MAACO, blocks_map
    import pyoracle as O
    g = blocks_map(512, 0.2, seed=4000)
    N = 4096
    dev = MAACO(g, N, 2, rng_seed=4, verbose=False, **MAACO_DEFAULT)
    orc = O.MaacoOracle(g, N, 2, seed=4, threads=0, **MAACO_DEFAULT)
    for it in (1, 2):
        _check_pass(dev, orc, it, N)
    nc, ln, tn, cells = dev.last_tours()
    C = 512
    ok = np.flatnonzero(nc > 0)
    assert ok.size > N // 4
    for a in ok[:256]:
        p = cells[a, :nc[a]]
        assert p[0] == 0 and p[-1] == 512 * 512 - 1
        assert len(set(p.tolist())) == len(p)                         # tabu: no revisits
        dr, dc = np.diff(p // C), np.diff(p % C)
        assert np.all(np.abs(dr) <= 1) and np.all(np.abs(dc) <= 1)    # 8-connected
        assert not np.any(g.ravel()[p] == 1)                          # never on an obstacle
        diag = (dr != 0) & (dc != 0)
        assert abs(ln[a] - (diag.sum() * np.sqrt(2.0) + (~diag).sum())) < 1e-9 * ln[a]
        # crossing prohibition MAACO.py:100-120
        r0, c0 = p[:-1] // C, p[:-1] % C
        assert not np.any(diag & ((g[r0 + dr, c0] == 1) | (g[r0, c0 + dc] == 1)))
    tau = dev.pheromone_matrix
    free = g != 1
    assert np.all(tau[~free] == 1e-9) and tau[free].min() > 0


def test_lane_layouts_agree():
    """Every packing of ants into warps is the same function."""
    from maaco_path_planing_b200 import MAACO, blocks_map
    g = blocks_map(128, 0.2, seed=31)
    outs = []
    for apw in (0, 1, 2, 8, 16, 32):
        dev = MAACO(g, 300, 2, rng_seed=5, ants_per_warp=apw, verbose=False, **MAACO_DEFAULT)
        dev.run_iteration(1)
        dev.run_iteration(2)
        outs.append((dev.last_tours(), dev.pheromone_matrix))
    for other in outs[1:]:
        for x, y in zip(outs[0][0], other[0]):
            assert np.array_equal(x, y)
        assert np.array_equal(outs[0][1], other[1])


def test_large_unaligned_map_vs_oracle():
    """1000 columns (not a multiple of 32: window words straddle bitmap words, write-back by atomics) and tours of
    ~1800 steps: dozens of window slides per ant in every direction."""
    from maaco_path_planing_b200 import MAACO, blocks_map
    import pyoracle as O
    g = blocks_map(1000, 0.2, seed=2003)
    N = 512
    dev = MAACO(g, N, 2, rng_seed=3, verbose=False, **MAACO_DEFAULT)
    orc = O.MaacoOracle(g, N, 2, seed=3, threads=0, **MAACO_DEFAULT)
    for it in (1, 2):
        _check_pass(dev, orc, it, N)
    nc = dev.last_tours()[0]
    assert (nc > 0).sum() > N // 2 and nc.max() > 1200


def test_ants_per_warp_env_override_agrees(monkeypatch):
    """The tour kernel gives the same colony for every packing of ants into warps (here forced through the
    MPP_TOUR_APW environment switch) on a map whose columns are not a multiple of the 32-cell tile."""
    from maaco_path_planing_b200 import MAACO, blocks_map
    import pyoracle as O
    g = blocks_map(150, 0.2, seed=77)                      # 150 columns: window words straddle bitmap words
    N = 200
    orc = O.MaacoOracle(g, N, 2, seed=9, threads=0, **MAACO_DEFAULT)
    ref = [orc.iterate(1), orc.iterate(2)]
    tau_ref = orc.tau.copy()
    for apw in (1, 4, 32, 0):
        if apw:
            monkeypatch.setenv("MPP_TOUR_APW", str(apw))
        else:
            monkeypatch.delenv("MPP_TOUR_APW", raising=False)
        dev = MAACO(g, N, 2, rng_seed=9, verbose=False, **MAACO_DEFAULT)
        for it in (1, 2):
            dev.run_iteration(it)
            nc, ln, tn, cells = dev.last_tours()
            ocells, onc, oln, otn, _ = ref[it - 1]
            assert np.array_equal(nc, onc) and np.array_equal(ln, oln) and np.array_equal(tn, otn)
            for a in range(N):
                assert np.array_equal(cells[a, :nc[a]], ocells[a, :onc[a]])
        assert np.array_equal(dev.pheromone_matrix.ravel(), tau_ref)


def test_errors_like_reference():
    from maaco_path_planing_b200 import MAACO
    g = np.zeros((5, 5), int)
    with pytest.raises(ValueError, match="MAACO: Start node not found."):
        MAACO(g, 4, 2, **MAACO_DEFAULT)
    g[0, 0] = 2
    with pytest.raises(ValueError, match="MAACO: Target node not found."):
        MAACO(g, 4, 2, **MAACO_DEFAULT)


def test_walled_in_start_fails_like_reference():
    from maaco_path_planing_b200 import MAACO
    g = np.zeros((6, 6), int)
    g[0, 0], g[5, 5] = 2, 3
    g[0, 1] = g[1, 0] = g[1, 1] = 1
    dev = MAACO(g, 8, 2, rng_seed=1, verbose=False, **MAACO_DEFAULT)
    path, length, turns = dev.solve_path_planning()
    assert path == [] and length == float("inf") and turns == float("inf")
    assert dev.convergence_curve_data == [None, None]


def test_sharded_colony_two_gpus():
    """2-GPU sharded colony == single-colony oracle (skipped on a 1-GPU box)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "multigpu_check ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_batched_maps_equal_individual_solves():
    """Config-5 style sweep: independent maps solved as waves of one launch per colony pass (grid.y = map) give
    exactly the per-map results -- every pass's per-ant records and pheromone fields, and the returned solutions."""
    from maaco_path_planing_b200 import MAACO, blocks_map
    from maaco_path_planing_b200.batch import MAACOBatch, solve_maaco_batch
    grids = [blocks_map(64, 0.2, seed=500 + i) for i in range(5)]
    seeds = [900 + i for i in range(5)]
    res = solve_maaco_batch(grids, 128, 4, MAACO_DEFAULT, seeds=seeds, wave=3)
    assert [r[0] for r in res] == list(range(5))
    solos = []
    for i, path, length, turns, curve in res:
        solo = MAACO(grids[i], 128, 4, rng_seed=seeds[i], verbose=False, **MAACO_DEFAULT)
        p2, l2, t2 = solo.solve_path_planning()
        assert path == p2 and length == l2 and turns == t2 and curve == solo.convergence_curve_data
        solos.append(solo)
    b = MAACOBatch(np.stack(grids), 128, 4, seeds=seeds, **MAACO_DEFAULT)
    refs = [MAACO(grids[i], 128, 4, rng_seed=seeds[i], verbose=False, **MAACO_DEFAULT) for i in range(5)]
    for it in (1, 2, 3):
        b.run_iteration(it)
        nc, ln, tn = b.last_results()
        tau = b.pheromone()
        for i, r in enumerate(refs):
            r.run_iteration(it)
            rn, rl, rt = r.last_results()
            assert np.array_equal(nc[i], rn) and np.array_equal(ln[i], rl) and np.array_equal(tn[i], rt)
            assert np.array_equal(tau[i], r.pheromone_matrix)
    # maps with different start / target cannot share the wave's tables: they are grouped apart
    g2 = [g.copy() for g in grids[:3]]
    g2[1][0, 0], g2[1][0, 5] = 0, 2
    res2 = solve_maaco_batch(g2, 64, 2, MAACO_DEFAULT, seeds=[1, 2, 3], wave=8)
    for i, path, length, turns, curve in res2:
        solo = MAACO(g2[i], 64, 2, rng_seed=[1, 2, 3][i], verbose=False, **MAACO_DEFAULT)
        assert (path, length, turns) == solo.solve_path_planning()


def test_pheromone_kernel_vs_numpy_and_tile_row_slices():
    """mpp_maaco_pheromone on synthetic slabs: (a) equals a plain sequential fold in ant order (MAACO.py:304-332) on
    every cell, (b) the update of a slice of tile rows from slice-shaped buffers (what a sharded colony's ranks run,
    with clear_slabs) writes exactly the cells of that slice with the same values."""
    import ctypes as C
    import torch
    from maaco_path_planing_b200 import MAACO, _lib, blocks_map
    g = blocks_map(0, 0.2, seed=77, rows=100, cols=150)                  # 4 x 5 tiles, ragged right / bottom edges
    R, Cc = g.shape
    N = 2500                                                              # > one list chunk (2048 ants)
    m = MAACO(g, N, 2, rng_seed=1, verbose=False, **MAACO_DEFAULT)
    dev = m.device
    TR, TC = m.tile_rows, m.tile_cols
    NW = (N + 31) // 32
    rng = np.random.default_rng(5)
    slabs = np.zeros((TR * TC, N, 32), np.uint32)
    touched = np.zeros((TR * TC, NW), np.uint32)
    pair = rng.random((TR * TC, N)) < 0.08
    pair[0, :] = True                                                     # the start tile: every ant
    for t, a in zip(*np.nonzero(pair)):
        rows = rng.random(32) < 0.4
        w = rng.integers(1, 2 ** 32, 32, dtype=np.uint64).astype(np.uint32) & rng.integers(0, 2 ** 32, 32, dtype=np.uint64).astype(np.uint32)
        slabs[t, a] = np.where(rows, w, 0)
        touched[t, a >> 5] |= np.uint32(1 << (a & 31))
    slabs[0, :, 0] |= 1                                                   # cell (0, 0): visited by all
    dep = rng.random(N) * 0.01
    dep[rng.random(N) < 0.3] = 0.0                                        # failed ants deposit nothing
    ok = np.zeros(NW, np.uint32)
    for a in np.flatnonzero(dep != 0.0):
        ok[a >> 5] |= np.uint32(1 << (a & 31))
    tau0 = rng.random(R * Cc) + 0.01
    st = _lib.MaacoState(123.5, 7, 0, 0, -1, 0.0, -1, -1)
    m._state.copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))
    m._deposit.copy_(torch.as_tensor(dep))
    m._okbits.copy_(torch.as_tensor(ok.view(np.int32)))
    L = _lib.lib()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    it = 3                                                                # parity 1
    # ---- numpy truth ----
    want = np.empty(R * Cc)
    tmax = (1.0 / (1.0 - 0.1)) * (1.0 / 123.5)
    tmin = tmax / (2.0 * max(R, Cc))
    for r in range(R):
        for c in range(Cc):
            t_ = (r >> 5) * TC + (c >> 5)
            t = tau0[r * Cc + c] * (1.0 - 0.1)
            col = (slabs[t_, :, r & 31] >> np.uint32(c & 31)) & 1
            for a in np.flatnonzero(col):
                t += dep[a]                                               # + 0.0 for non-depositing ants: exact
            want[r * Cc + c] = 1e-9 if g[r, c] == 1 else min(max(t, tmin), tmax)
    # ---- (a) whole map ----
    m._tau[:R * Cc].copy_(torch.as_tensor(tau0))
    sl = torch.as_tensor(slabs.view(np.int32).reshape(-1), device=dev)
    tb = torch.zeros(2 * TR * TC * NW, dtype=torch.int32, device=dev)
    tb[TR * TC * NW:].copy_(torch.as_tensor(touched.view(np.int32).reshape(-1)))      # parity it & 1 = 1
    tb[:TR * TC * NW] = -1                                                # the other parity must come back cleared
    _lib.check(L.mpp_maaco_pheromone(m._maps, C.byref(m._colony), _lib.ptr(sl), _lib.ptr(tb), N, 0, TR, 0.1, it, 0, None, 0, stream),
               "mpp_maaco_pheromone")
    got = m._tau[:R * Cc].cpu().numpy()
    assert np.array_equal(got, want)
    assert int(tb[:TR * TC * NW].abs().sum()) == 0
    assert np.array_equal(sl.cpu().numpy(), slabs.view(np.int32).reshape(-1))          # clear_slabs = 0: untouched
    # ---- (b) two slices of tile rows (the second one padded past the map, like the last rank's) ----
    m._tau[:R * Cc].copy_(torch.as_tensor(tau0))
    per = 3
    for row0 in (0, per):
        rows_in = max(0, min(TR, row0 + per) - row0)
        s_sl = np.zeros((per * TC, N, 32), np.uint32)
        s_tb = np.zeros((2, per * TC, NW), np.uint32)
        s_sl[:rows_in * TC] = slabs[row0 * TC:(row0 + rows_in) * TC]
        s_tb[it & 1, :rows_in * TC] = touched[row0 * TC:(row0 + rows_in) * TC]
        d_sl = torch.as_tensor(s_sl.view(np.int32).reshape(-1), device=dev)
        d_tb = torch.as_tensor(s_tb.view(np.int32).reshape(-1), device=dev)
        _lib.check(L.mpp_maaco_pheromone(m._maps, C.byref(m._colony), _lib.ptr(d_sl), _lib.ptr(d_tb), N, row0, per, 0.1, it, 1,
                                         None, 0, stream), "mpp_maaco_pheromone")
        # what was read (slabs of depositing ants) has been cleared
        left = d_sl.cpu().numpy().view(np.uint32).reshape(per * TC, N, 32)
        assert not left[:, dep != 0.0].any()
    assert np.array_equal(m._tau[:R * Cc].cpu().numpy(), want)


def test_host_buffer_pass_equals_device_pass():
    """MAACO.run_iteration_host (C ABI mpp_maaco_pass_host: host pheromone field in; per-ant records, best path and
    updated field out, synchronous) is the same function as the device-resident pass."""
    from maaco_path_planing_b200 import MAACO, blocks_map
    g = blocks_map(96, 0.2, seed=12)
    N, K = 256, 3
    a = MAACO(g, N, K, rng_seed=5, verbose=False, **MAACO_DEFAULT)
    b = MAACO(g, N, K, rng_seed=5, verbose=False, **MAACO_DEFAULT)
    tau = b.pheromone_matrix.ravel().copy()
    res = np.zeros(N, dtype=[("length", "<f8"), ("n_cells", "<i4"), ("turns", "<i4")])
    best = np.zeros(4096, np.int32)
    for it in range(1, K + 1):
        a.run_iteration(it)
        st = b.run_iteration_host(it, tau_in=tau, tau_out=tau, result_out=res, best_out=best)
        nc, ln, tn = a.last_results()
        assert np.array_equal(res["n_cells"], nc) and np.array_equal(res["length"], ln) and np.array_equal(res["turns"], tn)
        assert np.array_equal(tau, a.pheromone_matrix.ravel())
        sa = a._read_state()
        assert (st.best_len, st.best_turns, st.best_n_cells, st.best_ant) == (sa.best_len, sa.best_turns, sa.best_n_cells, sa.best_ant)
        assert np.array_equal(best[:st.best_n_cells], a._best_cells[:sa.best_n_cells].cpu().numpy())
    pa = a.solve_path_planning()
    assert [r * 96 + c for r, c in pa[0]] == best[:st.best_n_cells].tolist()
