"""The N>1 path on CPU: two gloo ranks shard a colony exactly the way the NCCL ranks do
(maaco_path_planing_b200/dist.py + the exchange protocol of csrc/mpp_maaco.cu): ONE all-gather of per-rank buffers
holding the per-ant results and the tours as move codes, replay of every ant into this rank's slice of tile rows,
slice-wise ordered deposit, all-gather of tau slices.  The tours and the arithmetic come from the oracle (there is no
GPU here); what is under test is the sharding / buffer layout / host-side sizing and that the ordered merge
reproduces the single-colony pheromone field bit-for-bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import MAACO_DEFAULT, ROOT

REC = np.dtype([("length", "<f8"), ("n_cells", "<i4"), ("turns", "<i4")])
DR = np.array([-1, -1, -1, 0, 0, 1, 1, 1])                                # move order MAACO.py:98
DC = np.array([-1, 0, 1, -1, 1, -1, 0, 1])


def pack_buffer(cells, ncell, length, turns, C, cap):
    """numpy statement of mpp_maaco_xpack: [n_local x mpp_ant_result][int32 total + 12 pad][codes]."""
    nl = len(ncell)
    hdr = 16 * nl + 16
    buf = np.zeros(hdr + cap, np.uint8)
    rec = np.zeros(nl, REC)
    rec["length"], rec["n_cells"], rec["turns"] = length, ncell, turns
    buf[:16 * nl] = rec.view(np.uint8)
    off = 0
    for a in range(nl):
        n = int(ncell[a])
        if n > 1:
            d = np.diff(cells[a, :n].astype(np.int64))
            dr = np.round(d / C).astype(int)
            dc = d - dr * C
            code = (dr + 1) * 3 + (dc + 1)
            code = code - (code > 4)
            if off + n - 1 <= cap:
                buf[hdr + off:hdr + off + n - 1] = code
            off += n - 1
    buf[16 * nl:16 * nl + 4] = np.array([off], np.int32).view(np.uint8)
    return buf


def unpack_visits(buf_all, world, nl, cap, start, R, C, row0, rows):
    """numpy statement of mpp_maaco_xunpack: records of all ants + the visited cells of every ant that fall into cell
    rows [row0, row0 + rows)."""
    hdr = 16 * nl + 16
    seg = hdr + cap
    recs, visits = [], []
    for s in range(world):
        b = buf_all[s * seg:(s + 1) * seg]
        rec = b[:16 * nl].view(REC)
        total = int(b[16 * nl:16 * nl + 4].view(np.int32)[0])
        assert total <= cap, "exchange buffer overflow"
        recs.append(rec)
        off = 0
        for a in range(nl):
            n = int(rec["n_cells"][a])
            mine = []
            if n > 0:
                codes = b[hdr + off:hdr + off + n - 1].astype(int)
                cell = start + np.concatenate([[0], np.cumsum(DR[codes] * C + DC[codes])])
                mine = [int(x) for x in cell if row0 <= x // C < row0 + rows]
                off += n - 1
            visits.append(mine)
    return np.concatenate(recs), visits


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pyoracle as O
        from maaco_path_planing_b200 import dist as dm
        from maaco_path_planing_b200.gridmap import blocks_map
        g = blocks_map(0, 0.2, seed=3, rows=70, cols=40)                   # 3 tile rows over 2 ranks: a padded slice
        R, C = g.shape
        N, K, seed = 48, 3, 17
        n = g.size
        lo, hi = dm.shard_range(N, world, rank)
        nl = hi - lo
        TR = (R + 31) // 32
        per = dm.padded_tile_rows(TR, world) // world                     # tile rows per rank
        npad = per * world * 32 * C
        full = O.MaacoOracle(g, N, K, seed=seed, **MAACO_DEFAULT)          # single-colony truth
        mine = O.MaacoOracle(g, N, K, seed=seed, **MAACO_DEFAULT)          # this rank's replica of tau
        grid = g.ravel()
        (sr, sc), _ = O.find_start_target(g)
        cap = 65536
        for it in range(1, K + 1):
            full.iterate(it)
            cells, ncell, length, turns, tabu = mine.tours(it, ant0=lo, n_ants=nl)     # only my ants
            buf = pack_buffer(cells, ncell, length, turns, C, cap)
            buf_all = torch.zeros(world * buf.size, dtype=torch.uint8)
            dm.exchange_buffers(buf_all, torch.from_numpy(buf), dist.group.WORLD)
            allrec, visits = unpack_visits(buf_all.numpy(), world, nl, cap, sr * C + sc, R, C, rank * per * 32, per * 32)
            # the replayed cells are exactly the visited set the oracle recorded (for my own ants: check)
            for a in range(nl):
                want = [c for c in cells[a, :ncell[a]] if rank * per * 32 <= c // C < (rank + 1) * per * 32]
                assert visits[lo + a] == [int(c) for c in want]
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
            # slice-wise ordered deposit == mpp_maaco_pheromone(tile_row0 = rank*per, buf_tile_rows = per)
            p = mine.p
            tau = mine.tau
            c0, c1 = rank * per * 32 * C, min(n, (rank + 1) * per * 32 * C)
            sl = np.zeros(per * 32 * C)
            acc = {cell: tau[cell] * (1.0 - p.rho) for cell in range(c0, c1)}
            for i in range(N):                                             # global ant order
                l = allrec["length"][i]
                if np.isfinite(l) and allrec["n_cells"][i] > 0 and l > 1e-6:
                    for cell in visits[i]:
                        acc[cell] += p.Q / l
            best = mine.best_len if np.isfinite(mine.best_len) else float(p.rows + p.cols)
            best = max(best, 1e-6)
            tmax = (1.0 / (1.0 - p.rho)) * (1.0 / best)
            tmin = tmax / (2.0 * max(p.cols, p.rows, 1))
            for cell in range(c0, c1):
                sl[cell - c0] = 1e-9 if grid[cell] == 1 else min(max(acc[cell], tmin), tmax)
            tau_full = torch.zeros(npad, dtype=torch.float64)
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
    assert dm.padded_tile_rows(16, 8) == 16 and dm.padded_tile_rows(3, 2) == 4 and dm.padded_tile_rows(1, 8) == 8
    from maaco_path_planing_b200.batch import shard_maps
    assert shard_maps(7) == (0, 7)
