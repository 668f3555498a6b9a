"""Batched sweeps over independent maps (BASELINE config 5: "10k independent 256x256 maps x population 1024,
MPA+MAACO, maps sharded across 8 B200").  Maps are independent units, so they are sharded over the ranks with no
data-path collective, and on each GPU a WAVE of same-shape maps is one mpp_map_batch: every colony pass of the
whole wave is four kernel launches (ranking, tours, best tracking, pheromone update with grid.y = map), enqueued
without host synchronisation.  Results are identical to solving each map alone with the same seed.

The eta'**beta / dist-to-target tables depend only on (shape, start, target, parameters), so a wave shares ONE
table computed on the host with libm (bit-identical to the reference), and tau0 of each map is that shared table
with the map's obstacles masked on the device (mpp_maaco_tables)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .gridmap import START_NODE_VAL, TARGET_NODE_VAL

INF = float("inf")


def shard_maps(n_maps, group=None):
    """[lo, hi) of the maps this rank owns."""
    if group is None:
        return 0, n_maps
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_maps + world - 1) // world
    return min(n_maps, rank * per), min(n_maps, (rank + 1) * per)


def _grids_u8(grids):
    """[n_maps, rows, cols] cell codes as uint8 (what mpp_map_batch_create takes): values clipped to 0..255 like
    np.clip(int grid), written straight into the byte array (an int64 wave of 128 256x256 maps is 67 MB: the
    intermediate int64 copies cost more host time than the wave's passes cost on the GPU)."""
    g = np.asarray(grids)
    if g.ndim != 3:
        raise ValueError("grids must be [n_maps, rows, cols]")
    if g.dtype == np.uint8:
        return np.ascontiguousarray(g)
    if g.dtype.kind not in "iub":
        g = g.astype(int)                                                     # (the reference indexes int grids)
    g8 = np.empty(g.shape, np.uint8)
    np.clip(g, 0, 255, out=g8, casting="unsafe")
    return g8


class MAACOBatch:
    """M same-shape maps that share start and target, one colony of `num_ants` ants per map, solved together.
    Same parameters as MAACO (MAACO.py:11-14); `seeds` = one Philox seed per map."""

    def __init__(self, grids, num_ants, num_iterations, alpha, beta, rho, Q, a_turn_coef, wh_max, wh_min, k_h_adaptive,
                 q0_initial, C0_initial_pheromone=0.1, *, seeds=None, device=None, max_cells=None, ants_per_warp=0,
                 reuse=None):
        import torch
        L = _lib.lib()
        g8 = _grids_u8(grids)
        self.n_maps, self.rows, self.cols = g8.shape
        flat = g8.reshape(self.n_maps, -1)
        if not (flat == START_NODE_VAL).any(axis=1).all():
            raise ValueError("MAACO: Start node not found.")                  # MAACO.py:35-36
        if not (flat == TARGET_NODE_VAL).any(axis=1).all():
            raise ValueError("MAACO: Target node not found.")                 # MAACO.py:37-38
        _lib.require_device()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        _lib.check(L.mpp_map_batch_create(g8.ctypes.data_as(C.c_void_p), self.n_maps, self.rows, self.cols,
                                          self.device_index, C.byref(h)), "mpp_map_batch_create")
        self._maps = h
        self.num_ants, self.num_iterations = int(num_ants), int(num_iterations)
        self.alpha, self.rho, self.Q = alpha, rho, Q
        self._params = _lib.MaacoParams(alpha, beta, rho, Q, a_turn_coef, wh_max, wh_min, k_h_adaptive, q0_initial,
                                        C0_initial_pheromone, num_iterations)
        M, R, Cc, N = self.n_maps, self.rows, self.cols, self.num_ants
        n = R * Cc
        TR = (R + 31) // 32
        if max_cells is None:
            max_cells = min(n, max(1024, 8 * (R + Cc)))
        self.max_cells = int(max_cells)
        self._apw = int(ants_per_warp)
        if seeds is None:
            seeds = [int.from_bytes(np.random.bytes(8), "little") for _ in range(M)]
        self.seeds = [int(s) & (2 ** 64 - 1) for s in seeds]
        dev = self.device
        f64, i32, i64, u8 = torch.float64, torch.int32, torch.int64, torch.uint8
        K = max(1, self.num_iterations)
        key = (M, R, Cc, N, self.max_cells, K, self.device_index)
        if reuse is not None and getattr(reuse, "_key", None) == key:
            # the previous wave's device buffers (same shapes): nothing to allocate, only `touched` / counters to clear
            for name in ("_tau", "_E01", "_dist_t", "_rank", "_slabs", "_touched", "_moves", "_result", "_deposit", "_okbits",
                         "_best_cells", "_steps", "_log"):
                setattr(self, name, getattr(reuse, name))
            self._touched.zero_()
            self._steps.zero_()
        else:
            self._tau = torch.empty((M, n), dtype=f64, device=dev)
            self._E01 = torch.empty(2 * n, dtype=f64, device=dev)             # shared by the whole wave
            self._dist_t = torch.empty(n, dtype=f64, device=dev)
            self._rank = torch.empty(M * L.mpp_maaco_rank_words(R, Cc), dtype=i32, device=dev)
            self._slabs = torch.empty(M * L.mpp_maaco_slab_words(TR, Cc, N), dtype=i32, device=dev)
            self._touched = torch.zeros(M * L.mpp_maaco_touched_words(TR, Cc, N), dtype=i32, device=dev)
            self._moves = torch.empty(M * N * self.max_cells, dtype=u8, device=dev)
            self._result = torch.zeros((M, N, 2), dtype=i64, device=dev)
            self._deposit = torch.zeros((M, N), dtype=f64, device=dev)
            self._okbits = torch.zeros((M, (N + 31) // 32), dtype=i32, device=dev)
            self._best_cells = torch.zeros((M, self.max_cells + 1), dtype=i32, device=dev)
            self._steps = torch.zeros(1, dtype=i64, device=dev)
            self._log = torch.zeros((M, K, 4), dtype=f64, device=dev)
        self._key = key
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.mpp_maaco_tables(self._maps, C.byref(self._params), _lib.ptr(self._tau), n, _lib.ptr(self._E01),
                                      _lib.ptr(self._dist_t), stream), "mpp_maaco_tables")
        self._seeds = torch.from_numpy(np.array(self.seeds, np.uint64).view(np.int64)).to(dev)
        st = bytes(_lib.MaacoState(INF, -1, 0, 0, -1, INF, -1, -1))
        self._state = torch.frombuffer(bytearray(st * M), dtype=u8).to(dev)
        self._colony = _lib.Colony(
            self._tau.data_ptr(), n, self._E01.data_ptr(), 0, self._rank.data_ptr(), self._slabs.data_ptr(),
            self._touched.data_ptr(), self._moves.data_ptr(), self.max_cells, K, self._result.data_ptr(),
            self._deposit.data_ptr(), self._okbits.data_ptr(), self._state.data_ptr(), self._best_cells.data_ptr(),
            self._log.data_ptr(), self._steps.data_ptr(), self._seeds.data_ptr(), None)
        self._iter_done = 0
        self.kernel_launches = 0

    def close(self):
        h, self._maps = getattr(self, "_maps", None), None
        if h:
            _lib.lib().mpp_map_batch_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_iteration(self, it):
        """Enqueue one colony pass of every map of the wave (asynchronous)."""
        import torch
        if not 1 <= it <= max(1, self.num_iterations):
            raise ValueError(f"iteration {it} outside 1..{self.num_iterations}")
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.lib().mpp_maaco_pass(self._maps, C.byref(self._colony), C.byref(self._params), it, self.num_ants,
                                             self._apw, stream), "mpp_maaco_pass")
        self.kernel_launches += 4
        self._iter_done = max(self._iter_done, it)

    def pheromone(self):
        return self._tau.cpu().numpy().reshape(self.n_maps, self.rows, self.cols)

    def last_results(self):
        """(n_cells, length, turns), each [n_maps, num_ants], of the last pass."""
        import torch
        torch.cuda.synchronize(self.device)
        rec = self._result.cpu().numpy().view(np.dtype([("length", "<f8"), ("n_cells", "<i4"), ("turns", "<i4")]))
        rec = rec.reshape(self.n_maps, self.num_ants)
        return rec["n_cells"].copy(), rec["length"].copy(), rec["turns"].copy()

    def total_steps(self):
        return int(self._steps.cpu().item())

    def solve(self):
        """Runs the remaining iterations; returns one (path, length, turns, convergence_curve) per map -- the values
        MAACO.solve_path_planning (MAACO.py:334-371) returns / stores for that map."""
        import torch
        for it in range(self._iter_done + 1, self.num_iterations + 1):
            self.run_iteration(it)
        torch.cuda.synchronize(self.device)
        M, K = self.n_maps, self.num_iterations
        st = np.frombuffer(self._state.cpu().numpy().tobytes(), dtype=np.dtype(
            [("best_len", "<f8"), ("best_turns", "<i4"), ("best_n_cells", "<i4"), ("best_iter", "<i4"), ("best_ant", "<i4"),
             ("iter_best_len", "<f8"), ("iter_best_turns", "<i4"), ("iter_best_ant", "<i4")]))
        if (st["best_n_cells"] - 1 > self.max_cells).any():
            raise _lib.MppError(f"a best path outgrew max_cells={self.max_cells}; re-run with a larger max_cells")
        best = self._best_cells.cpu().numpy()
        log = self._log.cpu().numpy()[:, :K]
        out = []
        for k in range(M):
            nb = int(st["best_n_cells"][k])
            pr, pc = np.divmod(best[k, :nb], self.cols)
            path = list(zip(pr.tolist(), pc.tolist()))
            turns = int(st["best_turns"][k]) if st["best_turns"][k] >= 0 else INF
            curve = [float(v) if v != INF else None for v in log[k, :, 2]]
            out.append((path, float(st["best_len"][k]), turns, curve))
        return out


def _waves(grids, lo, hi, wave):
    """Group this rank's maps into waves of same shape / start / target (what one mpp_map_batch needs)."""
    groups = {}
    for i in range(lo, hi):
        g = np.asarray(grids[i])
        s, t = np.argwhere(g == START_NODE_VAL), np.argwhere(g == TARGET_NODE_VAL)
        key = (g.shape, tuple(s[0]) if len(s) else None, tuple(t[0]) if len(t) else None)
        groups.setdefault(key, []).append(i)
    for idx in groups.values():
        for w0 in range(0, len(idx), wave):
            yield idx[w0:w0 + wave]


def solve_maaco_batch(grids, num_ants, num_iterations, params, seeds=None, wave=256, group=None, device=None,
                      max_cells=None):
    """Run MAACO on every grid of `grids` (this rank's shard when `group` is given), `wave` maps per launch.

    Returns a list of (map_index, path, length, turns, convergence_curve) for the maps of this rank; results
    are identical to solving each map alone (same seeds => same Philox streams)."""
    lo, hi = shard_maps(len(grids), group)
    out = []
    prev = None
    for idx in _waves(grids, lo, hi, wave):
        b = MAACOBatch(np.stack([np.asarray(grids[i]) for i in idx]), num_ants, num_iterations,
                       seeds=None if seeds is None else [seeds[i] for i in idx], device=device, max_cells=max_cells,
                       reuse=prev, **params)
        for i, (path, length, turns, curve) in zip(idx, b.solve()):
            out.append((i, path, length, turns, curve))
        b.close()
        prev = b
    out.sort(key=lambda r: r[0])
    return out


class MPABatch:
    """M same-shape maps, one MPA population of `num_predators` paths per map, every iteration of every map in one
    launch (mpp_mpa_iteration_batch); the stable sorts (MPA.py:333,412) are a batched device argsort whose index
    the kernel reads through, and the best-so-far cascade (MPA.py:415-437) is a few vector operations over the maps --
    no host round trip inside the solve.  Same parameters as MPA (MPA.py:10-18); `seeds` = one Philox seed per map."""

    def __init__(self, grids, num_predators, num_iterations, FADs_rate=0.2, P_const=0.5, levy_beta=1.5,
                 turn_penalty_factor=0.1, safety_penalty_factor=0.05, min_safe_distance=1.5, allow_diagonal_moves=True,
                 restrict_diagonal_near_obstacle=True, diagonal_obstacle_penalty=1000.0, *, seeds=None, device=None,
                 max_cells=None, heap_cap=None, warps_per_map=None):
        import math
        import torch
        from .engine import make_policy
        L = _lib.lib()
        g8 = _grids_u8(grids)
        self.n_maps, self.rows, self.cols = g8.shape
        flat = g8.reshape(self.n_maps, -1)
        if not (flat == START_NODE_VAL).any(axis=1).all():
            raise ValueError("MPA: Start node not found in grid.")            # MPA.py:36-37
        if not (flat == TARGET_NODE_VAL).any(axis=1).all():
            raise ValueError("MPA: Target node not found in grid.")           # MPA.py:38-39
        _lib.require_device()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self._grids8 = g8
        h = C.c_void_p()
        _lib.check(L.mpp_map_batch_create(self._grids8.ctypes.data_as(C.c_void_p), self.n_maps, self.rows, self.cols,
                                          self.device_index, C.byref(h)), "mpp_map_batch_create")
        self._maps = h
        self.num_predators, self.num_iterations = int(num_predators), int(num_iterations)
        self.FADs_rate, self.P_const, self.levy_beta = FADs_rate, P_const, levy_beta
        self.policy = make_policy(turn_penalty_factor, safety_penalty_factor, min_safe_distance, diagonal_obstacle_penalty,
                                  restrict_diagonal_near_obstacle, allow_diagonal_moves, mode=1)
        b = levy_beta                                                         # Levy sigma MPA.py:251-253
        num = math.gamma(1 + b) * math.sin(math.pi * b / 2)
        den = math.gamma((1 + b) / 2) * b * (2 ** ((b - 1) / 2))
        sig = (num / den) ** (1 / b) if den > 1e-9 else 1.0
        if isinstance(sig, complex):
            raise ValueError("levy_beta gives a complex Levy sigma in the reference expression (MPA.py:253)")
        self._levy_sigma = float(sig)
        n = self.rows * self.cols
        self.max_cells = int(max_cells or min(n, max(1024, 4 * (self.rows + self.cols))))
        self.heap_cap = int(heap_cap or min(8 * n, max(4096, n // 4)))
        if seeds is None:
            seeds = [int.from_bytes(np.random.bytes(8), "little") for _ in range(self.n_maps)]
        self.seeds = [int(s) & (2 ** 64 - 1) for s in seeds]
        sm = torch.cuda.get_device_properties(self.device).multi_processor_count
        # search slots (one warp each): three CTAs of 8 warps per SM in total, split evenly over the maps (at least one CTA each)
        self.warps_per_map = int(warps_per_map or max(8, ((sm * 24) // self.n_maps) // 8 * 8))
        self._ctor = dict(grids=grids, num_predators=num_predators, num_iterations=num_iterations, FADs_rate=FADs_rate,
                          P_const=P_const, levy_beta=levy_beta, turn_penalty_factor=turn_penalty_factor,
                          safety_penalty_factor=safety_penalty_factor, min_safe_distance=min_safe_distance,
                          allow_diagonal_moves=allow_diagonal_moves,
                          restrict_diagonal_near_obstacle=restrict_diagonal_near_obstacle,
                          diagonal_obstacle_penalty=diagonal_obstacle_penalty, seeds=self.seeds, device=device,
                          warps_per_map=warps_per_map)
        self.predator_evaluations = 0
        self.kernel_launches = 0

    def close(self):
        h, self._maps = getattr(self, "_maps", None), None
        if h:
            _lib.lib().mpp_map_batch_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self):
        """Returns one (path, length, turns, safety_p, diag_p, fitness, convergence_curve) per map -- what
        MPA.solve_path_planning (MPA.py:320-448) returns / stores for that map."""
        import torch
        t = torch
        L = _lib.lib()
        dev = self.device
        M, N, K, mc = self.n_maps, self.num_predators, self.num_iterations, self.max_cells
        words = (self.rows * self.cols + 31) // 32
        slots = self.warps_per_map * M
        slot_bytes = L.mpp_astar_slot_bytes(self.rows, self.cols, self.heap_cap)
        scratch = t.zeros(256 + max(slots, M) * slot_bytes, dtype=t.uint8, device=dev)
        cells = [t.zeros((M, N, mc), dtype=t.int32, device=dev) for _ in range(2)]
        ncell = [t.zeros((M, N), dtype=t.int32, device=dev) for _ in range(2)]
        stats = [t.zeros((M, N, 5), dtype=t.float64, device=dev) for _ in range(2)]
        tmp = t.empty((slots, mc), dtype=t.int32, device=dev)
        avoid = t.empty((slots, words), dtype=t.int32, device=dev)
        queue = t.zeros(M, dtype=t.int32, device=dev)
        status = t.zeros(1, dtype=t.int32, device=dev)
        counters = t.zeros(4, dtype=t.int64, device=dev)
        seeds = t.from_numpy(np.array(self.seeds, np.uint64).view(np.int64)).to(dev)
        stream = C.c_void_p(t.cuda.current_stream(dev).cuda_stream)
        # ---- initial population: one search per map, replicated (MPA.py:231-245) ----
        _lib.check(L.mpp_mpa_init_batch(self._maps, C.byref(self.policy), N, mc, _lib.ptr(cells[0]), _lib.ptr(ncell[0]),
                                        _lib.ptr(stats[0]), _lib.ptr(scratch), scratch.numel(), self.heap_cap,
                                        _lib.ptr(status), _lib.ptr(counters), stream), "mpp_mpa_init_batch")
        cells[0][:, 1:] = cells[0][:, :1]
        ncell[0][:, 1:] = ncell[0][:, :1]
        stats[0][:, 1:] = stats[0][:, :1]
        self.kernel_launches += 1
        rows = t.arange(M, device=dev)

        def elite(buf):
            order = t.sort(stats[buf][:, :, 4], dim=1, stable=True).indices.to(t.int32).contiguous()   # list.sort is stable
            e = order[:, 0].long()
            return order, stats[buf][rows, e], cells[buf][rows, e], ncell[buf][rows, e]

        order, best_stats, best_cells, best_n = elite(0)                       # MPA.py:322-330
        best_stats, best_cells, best_n = best_stats.clone(), best_cells.clone(), best_n.clone()
        curve = t.empty((K + 1, M), dtype=t.float64, device=dev)
        curve[0] = best_stats[:, 4]
        cur = 0
        for it in range(1, K + 1):
            ratio = it / K
            CF = 0.0 if ratio >= 1.0 else ((1.0 - ratio) ** (2.0 * ratio) if ratio > 0 else 1.0)   # MPA.py:336
            phase = 1 if it <= K / 3 else (2 if it <= 2 * K / 3 else 3)
            nxt = cur ^ 1
            _lib.check(L.mpp_mpa_iteration_batch(
                self._maps, C.byref(self.policy), N, it, phase, self.P_const, CF, self.FADs_rate, self._levy_sigma,
                self.levy_beta, _lib.ptr(seeds), _lib.ptr(order), _lib.ptr(cells[cur]), _lib.ptr(ncell[cur]),
                _lib.ptr(stats[cur]), mc, _lib.ptr(cells[nxt]), _lib.ptr(ncell[nxt]), _lib.ptr(stats[nxt]), _lib.ptr(tmp),
                _lib.ptr(avoid), _lib.ptr(scratch), scratch.numel(), self.warps_per_map, self.heap_cap, _lib.ptr(queue),
                _lib.ptr(status), _lib.ptr(counters), stream), "mpp_mpa_iteration_batch")
            self.kernel_launches += 1
            self.predator_evaluations += M * N
            cur = nxt
            order, cs, cc, cn = elite(cur)                                      # sort :412, population[0]
            # best-so-far cascade MPA.py:415-437, for every map at once (columns: length, turns, safety, diag, fitness)
            eq = lambda a, b: (a - b).abs() < 1e-9
            better = cs[:, 4] < best_stats[:, 4]
            tie = eq(cs[:, 4], best_stats[:, 4]) & ~better
            eL, eT, eS = eq(cs[:, 0], best_stats[:, 0]), eq(cs[:, 1], best_stats[:, 1]), eq(cs[:, 2], best_stats[:, 2])
            casc = (cs[:, 0] < best_stats[:, 0]) | (eL & (cs[:, 1] < best_stats[:, 1])) | \
                   (eL & eT & (cs[:, 2] < best_stats[:, 2])) | (eL & eT & eS & (cs[:, 3] < best_stats[:, 3]))
            upd = better | (tie & casc)
            best_stats = t.where(upd[:, None], cs, best_stats)
            best_cells = t.where(upd[:, None], cc, best_cells)
            best_n = t.where(upd, cn, best_n)
            curve[it] = best_stats[:, 4]
        t.cuda.synchronize(dev)
        st = int(status.item())
        if st:                                                                   # a heap or a path buffer overflowed: the solve
            kw = dict(self._ctor)                                                # is deterministic -- repeat it with more room
            if st == 1:
                if self.heap_cap >= 8 * self.rows * self.cols:
                    raise _lib.MppError("A* heap overflow at maximum capacity")
                kw.update(heap_cap=min(8 * self.rows * self.cols, self.heap_cap * 4), max_cells=self.max_cells)
            else:
                if self.max_cells >= 2 * self.rows * self.cols:
                    raise _lib.MppError("path buffer overflow at maximum capacity")
                kw.update(max_cells=min(2 * self.rows * self.cols, 2 * self.max_cells), heap_cap=self.heap_cap)
            again = MPABatch(**kw)
            again.predator_evaluations, again.kernel_launches = self.predator_evaluations, self.kernel_launches
            out = again.solve()
            self.predator_evaluations, self.kernel_launches = again.predator_evaluations, again.kernel_launches
            again.close()
            return out
        self.expansions = int(counters[0].item())
        bs, bc, bn, cv = best_stats.cpu().numpy(), best_cells.cpu().numpy(), best_n.cpu().numpy(), curve.cpu().numpy()
        out = []
        for k in range(M):
            n = int(bn[k])
            path = [(int(c) // self.cols, int(c) % self.cols) for c in bc[k, :n]]
            length = float(bs[k, 0]) if n > 1 else 0
            col = [float(v) if v != INF else None for v in cv[:, k]]
            out.append((path, length, int(bs[k, 1]), float(bs[k, 2]), float(bs[k, 3]), float(bs[k, 4]), col))
        return out


def solve_mpa_batch(grids, num_predators, num_iterations, params, seeds=None, wave=32, group=None, device=None):
    """MPA over this rank's shard of the maps, `wave` maps per launch (grouped by shape).  Returns
    (map_index, path, length, turns, safety_p, diag_p, fitness, convergence_curve) per map of this rank."""
    lo, hi = shard_maps(len(grids), group)
    out = []
    shapes = {}
    for i in range(lo, hi):
        shapes.setdefault(np.asarray(grids[i]).shape, []).append(i)
    for idx_all in shapes.values():
        for w0 in range(0, len(idx_all), wave):
            idx = idx_all[w0:w0 + wave]
            b = MPABatch(np.stack([np.asarray(grids[i]) for i in idx]), num_predators, num_iterations,
                         seeds=None if seeds is None else [seeds[i] for i in idx], device=device, **params)
            for i, r in zip(idx, b.solve()):
                out.append((i,) + tuple(r))
            b.close()
    out.sort(key=lambda r: r[0])
    return out
