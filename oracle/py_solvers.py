"""Oracle mirrors of the reference's PSO and GA solve loops -- TEST INFRASTRUCTURE ONLY.

Sequential Python control flow (exactly the reference's loop order, incl. PSO's asynchronous gbest) over
the C oracle's numeric functions (update / selection / breeding / connector fitness).  Pinned against
trajectories recorded from the unmodified reference (tests/golden/solver_cases.npz).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import pyoracle as O

INF = float("inf")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class PsoOracle:
    """pso.py:97-240."""

    def __init__(self, grid, num_iterations, num_particles, W, w, c1, c2, tpf, spf, msd, diag, seed):
        self.grid = np.asarray(grid)
        self.g8 = np.ascontiguousarray(self.grid, dtype=np.uint8)
        self.R, self.C = self.grid.shape
        self.K, self.N, self.W = num_iterations, num_particles, W
        self.w, self.c1, self.c2 = w, c1, c2
        self.pol = (tpf, spf, msd, diag)
        self.seed = seed
        self.max_vel = max(1.0, 0.15 * max(self.R, self.C))                      # pso.py:34
        self.curve = []

    def _fitness(self, wp_cells):
        cells, ncell, stats, _ = O.waypoint_fitness(self.grid, np.asarray(wp_cells, np.int32).reshape(1, -1), *self.pol)
        return cells[0, :ncell[0]].copy(), stats[0].copy()

    def _round(self, pos):                                                      # pso.py:61,69-70
        out = []
        for r, c in pos:
            ir, ic = int(round(r)), int(round(c))
            out.append(max(0, min(self.R - 1, ir)) * self.C + max(0, min(self.C - 1, ic)))
        return out

    def initialize(self):                                                       # pso.py:97-161
        N, W = self.N, self.W
        lo, hi = -self.max_vel / 5, self.max_vel / 5
        self.pos, self.vel, self.fit, self.path = [], [], [], []
        attempts = 0
        while len(self.pos) < N and attempts < N * 20:
            u = [O.stream_uniform(self.seed, 2, 0, attempts, d) for d in range(4 * W)]
            attempts += 1
            pos = [[0 + (self.R - 1 - 0) * u[2 * k], 0 + (self.C - 1 - 0) * u[2 * k + 1]] for k in range(W)]
            vel = [[lo + (hi - lo) * u[2 * W + 2 * k], lo + (hi - lo) * u[2 * W + 2 * k + 1]] for k in range(W)]
            path, st = self._fitness(self._round(pos))
            if len(path):
                self.pos.append(pos); self.vel.append(vel); self.fit.append(st[4]); self.path.append(path)
        assert self.pos, "no valid particle"
        gi = int(np.argmin(self.fit))                                           # first strictly best :121
        n0 = len(self.pos)
        while len(self.pos) < N:                                                # :159-160
            j = int(O.stream_uniform(self.seed, 3, 0, len(self.pos), 0) * len(self.pos))
            self.pos.append([p[:] for p in self.pos[j]]); self.vel.append([v[:] for v in self.vel[j]])
            self.fit.append(self.fit[j]); self.path.append(self.path[j])
        self.pos = np.array(self.pos); self.vel = np.array(self.vel)
        self.pbest_pos = self.pos.copy(); self.pbest_fit = np.array(self.fit); self.cur_fit = np.array(self.fit)
        self.gbest_pos = self.pos[gi].copy(); self.gbest_fit = self.fit[gi]; self.gbest_path = self.path[gi]
        self.n_unique = n0

    def solve(self):
        self.initialize()
        self.curve.append(self.gbest_fit)
        wp = np.zeros(self.W, np.int32)
        for it in range(self.K):                                                # pso.py:178-229
            for p in range(self.N):
                pos, vel = np.ascontiguousarray(self.pos[p]), np.ascontiguousarray(self.vel[p])
                O.lib().orc_pso_update_particle(_p(pos), _p(vel), _p(np.ascontiguousarray(self.pbest_pos[p])),
                                                _p(np.ascontiguousarray(self.gbest_pos)), self.W, C.c_double(self.w),
                                                C.c_double(self.c1), C.c_double(self.c2), C.c_double(self.max_vel),
                                                self.R, self.C, C.c_uint64(self.seed), it, p, _p(wp))
                self.pos[p], self.vel[p] = pos, vel
                path, st = self._fitness(wp)
                if len(path):
                    self.cur_fit[p] = st[4]
                    if st[4] < self.pbest_fit[p]:                               # :216
                        self.pbest_fit[p] = st[4]; self.pbest_pos[p] = pos
                        if st[4] < self.gbest_fit:                              # :222 (seen by later particles)
                            self.gbest_fit = st[4]; self.gbest_pos = pos.copy(); self.gbest_path = path
            self.curve.append(self.gbest_fit)
        return self.gbest_path, self.gbest_fit


class GaOracle:
    """ga_solver.py:95-223."""

    def __init__(self, grid, num_generations, population_size, W, mutation_rate, crossover_rate, tournament_size,
                 tpf, spf, msd, diag, seed):
        self.grid = np.asarray(grid)
        self.g8 = np.ascontiguousarray(self.grid, dtype=np.uint8)
        self.R, self.C = self.grid.shape
        self.K, self.N, self.W = num_generations, population_size, W
        self.mut, self.cx, self.k = mutation_rate, crossover_rate, tournament_size
        self.pol = (tpf, spf, msd, diag)
        self.seed = seed
        self.curve = []

    def _evaluate(self, chrom):
        cells, ncell, stats, _ = O.waypoint_fitness(self.grid, np.asarray(chrom, np.int32), *self.pol)
        return cells, ncell, stats

    def initialize(self):                                                       # ga_solver.py:95-133
        N, W = self.N, self.W
        pop = []
        attempts = 0
        while len(pop) < N and attempts < N * 20:
            d = 0
            chrom = []
            for _ in range(W):
                while True:
                    r = int(O.stream_uniform(self.seed, 5, 0, attempts, d) * self.R); d += 1
                    c = int(O.stream_uniform(self.seed, 5, 0, attempts, d) * self.C); d += 1
                    if self.grid[r, c] != 1:
                        break
                chrom.append(r * self.C + c)
            attempts += 1
            cells, ncell, stats = self._evaluate([chrom])
            if ncell[0] > 0:
                pop.append((chrom, stats[0].copy(), cells[0, :ncell[0]].copy()))
        assert pop
        while len(pop) < N:                                                     # :129-130
            j = int(O.stream_uniform(self.seed, 6, 0, len(pop), 0) * len(pop))
            pop.append(pop[j])
        pop.sort(key=lambda x: x[1][4])                                         # stable :132
        self.pop = pop

    def solve(self):
        self.initialize()
        best = self.pop[0]
        self.curve.append(best[1][4])
        N, W = self.N, self.W
        for gen in range(self.K):                                               # ga_solver.py:178-215
            fit = np.array([ind[1][4] for ind in self.pop])
            parents = [O.lib().orc_ga_select(_p(fit), N, self.k, C.c_uint64(self.seed), gen, t) for t in range(N)]
            new = []
            pair = 0
            while len(new) < N:
                p1, p2 = self.pop[parents[(2 * pair) % N]], self.pop[parents[(2 * pair + 1) % N]]
                c1, c2 = np.zeros(W, np.int32), np.zeros(W, np.int32)
                O.lib().orc_ga_breed_pair(_p(self.g8), self.R, self.C, _p(np.asarray(p1[0], np.int32)),
                                          _p(np.asarray(p2[0], np.int32)), W, C.c_double(self.cx), C.c_double(self.mut),
                                          C.c_uint64(self.seed), gen, pair, _p(c1), _p(c2))
                pair += 1
                for child in (c1, c2):
                    if len(new) >= N:
                        break
                    cells, ncell, stats = self._evaluate([child.tolist()])
                    if ncell[0] > 0:
                        new.append((child.tolist(), stats[0].copy(), cells[0, :ncell[0]].copy()))
                    else:
                        new.append(p1 if len(new) % 2 == 0 else p2)            # :204-205
            new.sort(key=lambda x: x[1][4])
            self.pop = new
            if self.pop[0][1][4] < best[1][4]:
                best = self.pop[0]
            self.curve.append(best[1][4])
        return best[2], best[1]
