"""MAACO -- drop-in for the reference class ``MAACO.MAACO`` (MAACO.py:10-377) whose colony pass
(tour construction, best tracking, pheromone update) runs as sm_100a kernels behind the C ABI.

Same constructor arguments, attributes and return values as the reference; additive keyword-only
arguments select the RNG stream seed, the device and the (optional) ``torch.distributed`` group
over which the colony is sharded.  No CPU fallback: without libmpp_b200.so or a B200 this raises.

Keyword-only additions (none changes a result; every combination is bit-identical, see tests/test_maaco_gpu.py):
  rng_seed        seed of the Philox streams; None = from os.urandom
  device, group   CUDA device index; torch.distributed group to shard the colony over
  exchange        "moves" (default) or "dense": how a sharded colony ships its tours between ranks
  use_rank        per-pass move-ranking tables (mpp_maaco_rank); False = literal selection rules at every step
  lanes_per_ant   tour kernel form: 0 = library default (one thread per ant), 8 / 16 / 32 = cooperative lanes
  ants_per_warp   packing hint for the thread-per-ant kernel when several colonies share the GPU (batch.py)
  max_cells       capacity of the per-ant path buffers (default 8*(rows+cols); the solve repeats itself with rows*cols
                  if the best path did not fit)
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _lib
from . import dist as dist_mod
from .dist import padded_words
from .gridmap import GridMap, START_NODE_VAL, TARGET_NODE_VAL

INF = float("inf")


def _fresh_seed():
    """Seed of a solver built without rng_seed=: fresh entropy, like the reference's never-seeded RNGs -- unless
    MPP_RNG_SEED is set, which makes unmodified callers (the reference's main.py) reproducible."""
    env = os.environ.get("MPP_RNG_SEED")
    if env:
        return int(env, 0)
    return int.from_bytes(os.urandom(8), "little")


class MAACO:
    def __init__(self, grid, num_ants, num_iterations,
                 alpha, beta, rho, Q,
                 a_turn_coef, wh_max, wh_min, k_h_adaptive, q0_initial,
                 C0_initial_pheromone=0.1, *,
                 rng_seed=None, device=None, max_cells=None, lanes_per_ant=0, ants_per_warp=0, group=None,
                 exchange="moves",
                 use_rank=True, verbose=True):
        import torch
        self._ctor = dict(grid=grid, num_ants=num_ants, num_iterations=num_iterations, alpha=alpha, beta=beta, rho=rho,
                          Q=Q, a_turn_coef=a_turn_coef, wh_max=wh_max, wh_min=wh_min, k_h_adaptive=k_h_adaptive,
                          q0_initial=q0_initial, C0_initial_pheromone=C0_initial_pheromone, rng_seed=rng_seed,
                          device=device, lanes_per_ant=lanes_per_ant, ants_per_warp=ants_per_warp, group=group,
                          exchange=exchange, use_rank=use_rank, verbose=verbose)
        self.grid = np.array(grid, dtype=int)                       # MAACO.py:15
        self.rows, self.cols = self.grid.shape
        self.num_ants = num_ants
        self.num_iterations = num_iterations
        self.alpha, self.beta, self.rho, self.Q = alpha, beta, rho, Q
        self.a_turn_coef, self.wh_max, self.wh_min = a_turn_coef, wh_max, wh_min
        self.k_h_adaptive, self.q0_initial = k_h_adaptive, q0_initial
        self.k0_iter_threshold_factor = 0.7
        self.C0_base = C0_initial_pheromone
        s = np.argwhere(self.grid == START_NODE_VAL)
        t = np.argwhere(self.grid == TARGET_NODE_VAL)
        if not s.size > 0:
            raise ValueError("MAACO: Start node not found.")       # MAACO.py:35-36
        if not t.size > 0:
            raise ValueError("MAACO: Target node not found.")      # MAACO.py:37-38
        self.start_node = (int(s[0][0]), int(s[0][1]))
        self.target_node = (int(t[0][0]), int(t[0][1]))
        d = math.sqrt((self.start_node[0] - self.target_node[0]) ** 2 + (self.start_node[1] - self.target_node[1]) ** 2)
        self.dist_S_to_T_overall = d if d >= 1e-9 else 1e-9
        self.rng_seed = _fresh_seed() if rng_seed is None else int(rng_seed)
        self._ctor["rng_seed"] = self.rng_seed
        self.verbose = verbose
        # ants_per_warp: hint for the thread-per-ant kernel when several colonies share the GPU (batch.py)
        self.lanes_per_ant = -int(ants_per_warp) if ants_per_warp and lanes_per_ant in (0, 1) else lanes_per_ant
        if exchange not in ("moves", "dense"):
            raise ValueError("exchange must be 'moves' or 'dense'")
        self.exchange = exchange

        # ---- sharding of the colony over the process group (ants are independent given tau) ----
        self.group = group
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        if num_ants % self.world:
            raise ValueError("num_ants must be divisible by the group size")
        self.n_local = num_ants // self.world
        self.ant_offset = self.rank * self.n_local

        L = _lib.lib()
        self.map = GridMap(self.grid, device=device)
        self.device = torch.device("cuda", self.map.device)
        n = self.rows * self.cols
        self.n_words = padded_words(n, self.world)                  # bitmap words per ant (padded to split evenly)
        self.words_per_rank = self.n_words // self.world
        # per-ant path capacity: tours are a few (R+C) cells long (no backtracking, MAACO.py:278-302); a tour that
        # outgrows the buffer is still constructed and counted exactly -- only its cell list is truncated -- and
        # solve_path_planning() re-runs the (deterministic) solve with full capacity if the best path was cut
        self._auto_cells = max_cells is None
        if max_cells is None:
            max_cells = min(n, max(1024, 8 * (self.rows + self.cols)))
        self.max_cells = int(max_cells)
        dev = self.device
        f64, i32, i64 = torch.float64, torch.int32, torch.int64
        npad = self.n_words * 32
        self._tau = torch.zeros(npad, dtype=f64, device=dev)        # padded so tau slices all-gather evenly
        self._E01 = torch.empty(2 * n, dtype=f64, device=dev)       # eta'**beta, interleaved by turn flag
        self._dist_t = torch.empty(n, dtype=f64, device=dev)
        self.use_rank = use_rank
        self._rank = (torch.empty(_lib.lib().mpp_maaco_rank_words(self.map.handle), dtype=i32, device=dev)
                      if use_rank else None)
        self._params = _lib.MaacoParams(alpha, beta, rho, Q, a_turn_coef, wh_max, wh_min, k_h_adaptive, q0_initial,
                                        C0_initial_pheromone, num_iterations)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.mpp_maaco_tables(self.map.handle, C.byref(self._params), _lib.ptr(self._tau), _lib.ptr(self._E01),
                                      _lib.ptr(self._dist_t), C.c_void_p(stream)),
                   "mpp_maaco_tables")
        # word-major visited bitmaps of the local ants: [n_words][n_local]
        self._visit_local = torch.zeros(self.n_words * self.n_local, dtype=i32, device=dev)
        if self.world > 1:
            self._visit_recv = torch.empty(self.n_words * self.n_local, dtype=i32, device=dev)  # [G][Wn][n_local]
            self._side = torch.cuda.Stream(device=dev)
            if self.exchange == "moves":
                if self.cols < 3:
                    raise ValueError("the move-code exchange needs at least 3 columns")
                self._visit_recv.zero_()                              # rebuilt per pass, cleared by the update
                self._offsets = torch.zeros(num_ants, dtype=i32, device=dev)
                self._totals = torch.zeros(self.world, dtype=i32, device=dev)
                self._totals_host = torch.zeros(self.world, dtype=i32).pin_memory()
                self._xstatus = torch.zeros(1, dtype=i32, device=dev)
                self._packed_cap = 0
        self._cells = torch.zeros(self.n_local * self.max_cells, dtype=i32, device=dev)
        self._result = torch.zeros((num_ants, 2), dtype=i64, device=dev)   # mpp_ant_result per global ant
        self._deposit = torch.zeros(num_ants, dtype=f64, device=dev)
        self._best_cells = torch.zeros(self.max_cells, dtype=i32, device=dev)
        self._steps = torch.zeros(1, dtype=i64, device=dev)
        self._log = torch.zeros(max(1, num_iterations) * 4, dtype=f64, device=dev)
        st = _lib.MaacoState(INF, -1, 0, 0, -1, INF, -1, -1)
        self._state = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8).to(dev)

        self.best_path_overall = []
        self.best_path_length_overall = INF
        self.best_path_turns_overall = INF
        self.convergence_curve_data = []
        self._iter_done = 0
        self.kernel_launches = 0

    # ---- reference attributes materialised from device state ------------------------------
    @property
    def pheromone_matrix(self):
        n = self.rows * self.cols
        return self._tau[:n].cpu().numpy().reshape(self.rows, self.cols)

    @property
    def dist_to_target_matrix(self):
        return self._dist_t.cpu().numpy().reshape(self.rows, self.cols)

    def _calculate_adaptive_q0(self, current_iteration_num):
        return _lib.lib().mpp_maaco_q0(self.num_iterations, int(current_iteration_num), self.q0_initial)

    # ---- one colony pass (MAACO.py:336-359), fully asynchronous ----------------------------
    def _enqueue_tours(self, it, stream, mid_event=None):
        nl, off = self.n_local, self.ant_offset
        if self.use_rank:                                            # move ranking for the current tau
            _lib.check(_lib.lib().mpp_maaco_rank(self.map.handle, _lib.ptr(self._tau), _lib.ptr(self._E01), self.alpha,
                                                 _lib.ptr(self._rank), stream), "mpp_maaco_rank")
        if mid_event is not None:
            mid_event.record()
        _lib.check(_lib.lib().mpp_maaco_tours(
            self.map.handle, _lib.ptr(self._tau), _lib.ptr(self._E01), _lib.ptr(self._rank) if self.use_rank else None, it,
            self._calculate_adaptive_q0(it), self.alpha, nl, off, C.c_uint64(self.rng_seed),
            _lib.ptr(self._visit_local), _lib.ptr(self._cells), self.max_cells,
            C.c_void_p(self._result.data_ptr() + 16 * off), _lib.ptr(self._steps), self.lanes_per_ant, stream),
            "mpp_maaco_tours")

    def _enqueue_best(self, it, stream):
        _lib.check(_lib.lib().mpp_maaco_best(
            _lib.ptr(self._result), _lib.ptr(self._cells), self.max_cells, self.ant_offset, self.n_local,
            self.num_ants, self.Q, it, _lib.ptr(self._state), _lib.ptr(self._best_cells), _lib.ptr(self._deposit),
            _lib.ptr(self._log), stream), "mpp_maaco_best")

    def _enqueue_pheromone(self, stream):
        L = _lib.lib()
        if self.world == 1:
            _lib.check(L.mpp_maaco_pheromone(self.map.handle, _lib.ptr(self._tau), _lib.ptr(self._visit_local),
                                             _lib.ptr(self._deposit), 1, self.num_ants, 0, self.n_words, self.rho,
                                             _lib.ptr(self._state), 1, stream), "mpp_maaco_pheromone")
        else:
            wn = self.words_per_rank
            _lib.check(L.mpp_maaco_pheromone(self.map.handle, _lib.ptr(self._tau), _lib.ptr(self._visit_recv),
                                             _lib.ptr(self._deposit), self.world, self.n_local, self.rank * wn, wn,
                                             self.rho, _lib.ptr(self._state), 1 if self.exchange == "moves" else 0,
                                             stream), "mpp_maaco_pheromone")

    def _enqueue_iteration(self, it, events=None):
        """One colony pass.  `events`: optional CUDA events recorded at (start, after tours, before the
        pheromone update, end[, after the ranking kernel]) on the launching stream -- used by bench.py for
        per-kernel timing."""
        import torch
        if not 1 <= it <= max(1, self.num_iterations):               # the per-iteration log has num_iterations rows
            raise ValueError(f"iteration {it} outside 1..{self.num_iterations}")
        cur = torch.cuda.current_stream(self.device)
        stream = C.c_void_p(cur.cuda_stream)
        if events:
            events[0].record(cur)
        self._enqueue_tours(it, stream, events[4] if events and len(events) > 4 else None)
        if events:
            events[1].record(cur)
        if self.world > 1:
            nl, off = self.n_local, self.ant_offset
            L = _lib.lib()
            # in-place all-gather: this rank's slice of the result table is already in position
            dist_mod.exchange_results(self._result, self._result[off:off + nl], self.group)
            if self.exchange == "dense":
                dist_mod.exchange_visit_slices(self._visit_recv, self._visit_local, self.group)
            else:
                _lib.check(L.mpp_maaco_move_offsets(_lib.ptr(self._result), self.world, nl, _lib.ptr(self._offsets),
                                                    _lib.ptr(self._totals), stream), "mpp_maaco_move_offsets")
                self._totals_host.copy_(self._totals, non_blocking=True)
                cur.synchronize()                                      # the all-gather below is sized on the host
                cap = ((int(self._totals_host.max()) + 65535) // 65536) * 65536
                if cap > self._packed_cap:
                    self._packed_cap = cap
                    self._packed_local = torch.empty(cap, dtype=torch.uint8, device=self.device)
                    self._packed_all = torch.empty(cap * self.world, dtype=torch.uint8, device=self.device)
                cap = self._packed_cap
                _lib.check(L.mpp_maaco_pack_moves(self.map.handle, _lib.ptr(self._cells), self.max_cells,
                                                  C.c_void_p(self._result.data_ptr() + 16 * off),
                                                  C.c_void_p(self._offsets.data_ptr() + 4 * off), nl,
                                                  _lib.ptr(self._packed_local), cap, _lib.ptr(self._xstatus), stream),
                           "mpp_maaco_pack_moves")
                dist_mod.exchange_moves(self._packed_all, self._packed_local, self.group)
                wn = self.words_per_rank
                _lib.check(L.mpp_maaco_rebuild_visits(self.map.handle, _lib.ptr(self._packed_all), cap,
                                                      _lib.ptr(self._offsets), _lib.ptr(self._result), self.world, nl,
                                                      self.rank * wn, wn, _lib.ptr(self._visit_recv), stream),
                           "mpp_maaco_rebuild_visits")
                self.kernel_launches += 3
            # the local bitmaps are free again: clear them off the critical path
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self._visit_local.zero_()
        self._enqueue_best(it, stream)
        if events:
            events[2].record(cur)
        self._enqueue_pheromone(stream)
        if self.world > 1:
            wn32 = self.words_per_rank * 32
            dist_mod.gather_tau(self._tau, self._tau[self.rank * wn32:(self.rank + 1) * wn32], self.group)
            cur.wait_stream(self._side)                                # next pass's tours need the cleared bitmaps
        if events:
            events[3].record(cur)
        self.kernel_launches += 4 if self.use_rank else 3

    def _read_state(self):
        st = _lib.MaacoState.from_buffer_copy(self._state.cpu().numpy().tobytes())
        return st

    def solve_path_planning(self):
        import torch
        K = self.num_iterations
        for it in range(self._iter_done + 1, K + 1):
            self._enqueue_iteration(it)
        torch.cuda.synchronize(self.device)
        self._iter_done = K
        st = self._read_state()
        overflow = st.best_n_cells > self.max_cells
        if self.world > 1:
            import torch.distributed as dist
            # every rank must take the same decision: the flag of the move-code exchange is rank-local
            flag = torch.tensor([int(overflow) | (int(self._xstatus.item()) if self.exchange == "moves" else 0)],
                                dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            overflow = bool(flag.item())
        if overflow:
            if not self._auto_cells or self.max_cells >= self.rows * self.cols:
                raise _lib.MppError(f"a tour outgrew max_cells={self.max_cells} (best path: {st.best_n_cells} cells); "
                                    "re-run with a larger max_cells")
            # the solve is a deterministic function of (grid, parameters, seed): repeat it with full path capacity
            again = MAACO(max_cells=self.rows * self.cols, **self._ctor)
            launches = self.kernel_launches + again.kernel_launches
            self.__dict__.update(again.__dict__)
            self.kernel_launches = launches
            return self.solve_path_planning()
        log = self._log.cpu().numpy().reshape(-1, 4)[:K]
        if self.world > 1 and st.best_n_cells > 0:
            import torch.distributed as dist
            owner = st.best_ant // self.n_local                     # rank that constructed the best ant
            dist.broadcast(self._best_cells, src=dist.get_global_rank(self.group, owner), group=self.group)
            torch.cuda.synchronize(self.device)
        cells = self._best_cells[:st.best_n_cells].cpu().numpy()
        self.best_path_overall = [(int(c) // self.cols, int(c) % self.cols) for c in cells]
        self.best_path_length_overall = float(st.best_len)
        self.best_path_turns_overall = int(st.best_turns) if st.best_turns >= 0 else INF
        self.convergence_curve_data = [float(r[2]) if r[2] != INF else None for r in log]  # MAACO.py:360-362
        if self.verbose and self.rank == 0:
            tfmt = lambda v: int(v) if v >= 0 else INF
            for it in range(1, K + 1):                               # MAACO.py:363-366 (printed after the run)
                if it % 10 == 0 or it == 1 or it == K:
                    r = log[it - 1]
                    print(f"MAACO Iter {it}/{K}: Iter Best L={r[0]:.2f}, T={tfmt(r[1])}, "
                          f"Overall Best L={r[2]:.2f}, T={tfmt(r[3])}")
            if self.best_path_overall:
                print(f"\nMAACO Solved: Length={self.best_path_length_overall:.2f}, "
                      f"Turns={self.best_path_turns_overall}")
            else:
                print("\nMAACO: No solution found.")
        return self.best_path_overall, self.best_path_length_overall, self.best_path_turns_overall

    # ---- per-iteration access used by the parity tests and the benchmark -------------------
    def run_iteration(self, it):
        """Enqueue one colony pass; returns nothing (read results with `last_tours`)."""
        self._enqueue_iteration(it)
        self._iter_done = max(self._iter_done, it)

    def last_results(self):
        """(n_cells, length, turns) of every ant of the colony in the last pass (numpy)."""
        import torch
        torch.cuda.synchronize(self.device)
        raw = self._result.cpu().numpy()
        rec = raw.view(np.dtype([("length", "<f8"), ("n_cells", "<i4"), ("turns", "<i4")])).reshape(-1)
        return rec["n_cells"].copy(), rec["length"].copy(), rec["turns"].copy()

    def last_tours(self):
        """(n_cells, length, turns, cells[n_local, <= max_cells]) of this rank's ants in the last pass (only the
        columns some tour reached are copied to the host)."""
        nc, ln, tn = self.last_results()
        sl = slice(self.ant_offset, self.ant_offset + self.n_local)
        used = int(min(self.max_cells, max(1, nc[sl].max(initial=1))))
        cells = self._cells.view(self.n_local, self.max_cells)[:, :used].cpu().numpy()
        return nc[sl], ln[sl], tn[sl], cells

    def total_steps(self):
        return int(self._steps.cpu().item())

    # ---- plotting hooks of the reference (MAACO.py:373-377): out of scope, forwarded if possible --
    def visualize_pheromone_matrix(self, title="Mức Pheromone MAACO"):
        try:
            from visualization import visualize_pheromone_matrix as viz
        except Exception:
            print("visualization/matplotlib not available; pheromone_matrix is exposed as an ndarray")
            return
        viz(self.grid, self.pheromone_matrix, title)

    def plot_convergence_curve(self):
        try:
            from visualization import plot_convergence_curve as viz
        except Exception:
            print("visualization/matplotlib not available; convergence_curve_data is exposed as a list")
            return
        viz(self.convergence_curve_data, "MAACO", color='orangered')
