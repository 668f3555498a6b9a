"""The N>1 path on CPU: two gloo ranks shard a colony exactly the way the NCCL ranks do
(maaco_path_planing_b200/dist.py): all-gather of per-ant results, all-to-all of visited-bitmap word
slices, slice-wise ordered deposit, all-gather of tau slices.  The tours and the arithmetic come from
the oracle (there is no GPU here); what is under test is the sharding / exchange layout and that the
ordered merge reproduces the single-colony pheromone field bit-for-bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import MAACO_DEFAULT, ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pyoracle as O
        from maaco_path_planing_b200 import dist as dm
        from maaco_path_planing_b200.gridmap import blocks_map
        g = blocks_map(40, 0.2, seed=3)
        N, K, seed = 48, 3, 17
        n = g.size
        lo, hi = dm.shard_range(N, world, rank)
        nl = hi - lo
        W = dm.padded_words(n, world)
        wn = W // world
        full = O.MaacoOracle(g, N, K, seed=seed, **MAACO_DEFAULT)          # single-colony truth
        mine = O.MaacoOracle(g, N, K, seed=seed, **MAACO_DEFAULT)          # this rank's replica of tau
        grid = g.ravel()
        for it in range(1, K + 1):
            full.iterate(it)
            cells, ncell, length, turns, tabu = mine.tours(it, ant0=lo, n_ants=nl)     # only my ants
            # per-ant results, packed like mpp_ant_result (length f64, n_cells i32, turns i32)
            rec = np.zeros(nl, dtype=[("length", "<f8"), ("n_cells", "<i4"), ("turns", "<i4")])
            rec["length"], rec["n_cells"], rec["turns"] = length, ncell, turns
            res_all = torch.zeros((N, 2), dtype=torch.int64)
            dm.exchange_results(res_all, torch.from_numpy(rec.view(np.int64).reshape(nl, 2).copy()), dist.group.WORLD)
            allrec = res_all.numpy().view(rec.dtype).reshape(-1)
            # word-major visited bitmaps of my ants: [W][nl]
            vis = np.zeros((W, nl), np.int32)
            vis[:tabu.shape[1], :] = tabu.view(np.int32).T
            recv = torch.zeros(W * nl, dtype=torch.int32)
            dm.exchange_visit_slices(recv, torch.from_numpy(vis.reshape(-1).copy()), dist.group.WORLD)
            seg = recv.numpy().view(np.uint32).reshape(world, wn, nl)       # [source rank][my words][its ants]
            # replicated best tracking (MAACO.py:343-358) on the gathered results
            bl, bt, bi = np.inf, -1, -1
            for i in range(N):
                l, t_ = allrec["length"][i], allrec["turns"][i]
                if l < bl:
                    bl, bt, bi = l, t_, i
                elif abs(l - bl) < 1e-9 and t_ >= 0 and (bt < 0 or t_ < bt):
                    bt, bi = t_, i
            if bl < mine.best_len:
                mine.best_len, mine.best_turns = bl, bt
            elif abs(bl - mine.best_len) < 1e-9 and bt < mine.best_turns:
                mine.best_turns = bt
            # slice-wise ordered deposit == mpp_maaco_pheromone(n_seg=world, seg_ants=nl, word0=rank*wn, n_words=wn)
            p = mine.p
            tau = mine.tau
            c0, c1 = rank * wn * 32, min(n, (rank + 1) * wn * 32)
            sl = np.zeros(wn * 32)
            for cell in range(c0, c1):
                t = tau[cell] * (1.0 - p.rho)
                w, b = divmod(cell - c0, 32)
                for s in range(world):
                    for a in range(nl):
                        if (seg[s, w, a] >> b) & 1:
                            l = allrec["length"][s * nl + a]
                            if np.isfinite(l) and allrec["n_cells"][s * nl + a] > 0 and l > 1e-6:
                                t += p.Q / l
                best = mine.best_len if np.isfinite(mine.best_len) else float(p.rows + p.cols)
                best = max(best, 1e-6)
                tmax = (1.0 / (1.0 - p.rho)) * (1.0 / best)
                tmin = tmax / (2.0 * max(p.cols, p.rows, 1))
                sl[cell - c0] = 1e-9 if grid[cell] == 1 else min(max(t, tmin), tmax)
            tau_full = torch.zeros(W * 32, dtype=torch.float64)
            dm.gather_tau(tau_full, torch.from_numpy(sl), dist.group.WORLD)
            mine.tau[:] = tau_full.numpy()[:n]
            assert np.array_equal(mine.tau, full.tau), f"rank {rank}: tau differs after pass {it}"
            assert mine.best_len == full.best_len and mine.best_turns == full.best_turns
        # independent maps (config 5) are sharded by contiguous index ranges that cover every map exactly once
        from maaco_path_planing_b200.batch import shard_maps
        for n_maps in (1, 5, 48):
            lo_m, hi_m = shard_maps(n_maps, dist.group.WORLD)
            owned = torch.zeros(n_maps, dtype=torch.int64)
            owned[lo_m:hi_m] = 1
            dist.all_reduce(owned)
            assert bool((owned == 1).all()), f"maps not covered exactly once: {owned.tolist()}"
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_colony_exchange_two_gloo_ranks():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_helpers():
    from maaco_path_planing_b200 import dist as dm
    assert dm.shard_range(4096, 8, 3) == (1536, 2048)
    with pytest.raises(ValueError):
        dm.shard_range(10, 4, 0)
    assert dm.padded_words(512 * 512, 8) == 8192 and dm.padded_words(20 * 20, 8) == 16
    from maaco_path_planing_b200.batch import shard_maps
    assert shard_maps(7) == (0, 7)
