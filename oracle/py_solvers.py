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


class _Stream:
    """Cursor over one Philox stream with CPython's derived draws (randbelow = floor(u*n), Kinderman-Monahan)."""
    NV_MAGICCONST = 1.7155277699214135

    def __init__(self, seed, cls, it, ind):
        self.k = (seed, cls, it, ind)
        self.d = 0

    def random(self):
        u = O.stream_uniform(*self.k, self.d)
        self.d += 1
        return u

    def below(self, n):
        j = int(self.random() * n)
        return j if j < n else n - 1

    def randint(self, a, b):
        return a + self.below(b - a + 1)

    def uniform(self, a, b):
        return a + (b - a) * self.random()

    def normalvariate(self, mu, sigma):
        import math
        while True:
            u1 = self.random()
            u2 = 1.0 - self.random()
            z = self.NV_MAGICCONST * (u1 - 0.5) / u2
            if z * z / 4.0 <= -math.log(u2):
                break
        return mu + z * sigma


class MpaOracle:
    """MPA.py:231-448 (population of paths; phases, memory, FADs, sorts, best cascade)."""

    def __init__(self, grid, num_predators, num_iterations, FADs_rate, P_const, levy_beta, tpf, spf, msd, diag, seed):
        import math
        self.grid = np.asarray(grid)
        self.R, self.C = self.grid.shape
        self.N, self.K = num_predators, num_iterations
        self.FADs, self.P, self.beta = FADs_rate, P_const, levy_beta
        self.pol = (tpf, spf, msd, diag)
        self.seed = seed
        (sr, sc), (tr, tc) = O.find_start_target(self.grid)
        self.S, self.T = sr * self.C + sc, tr * self.C + tc
        self.astar = O.AStarOracle(self.grid, True, True)
        b = levy_beta                                                           # MPA.py:251-253
        num = math.gamma(1 + b) * math.sin(math.pi * b / 2)
        den = math.gamma((1 + b) / 2) * b * (2 ** ((b - 1) / 2))
        self.sigma = (num / den) ** (1 / b) if den > 1e-9 else 1.0
        self.curve = []

    def _a_star(self, src, dst, avoid=()):                                      # MPA.py:106-151
        bits = O.cells_to_bits(list(avoid), self.grid.size) if avoid else None
        cells, g, _, _ = self.astar.solve(1, src, dst, bits)
        return list(cells)

    def _stats(self, path):                                                     # MPA.py:215-229
        return O.path_stats(self.grid, path, self.pol[0], self.pol[1], self.pol[2], self.pol[3], True, mode=1)

    def _ind(self, path):
        return {"path": list(path), "stats": self._stats(path)}

    def _free(self, cell):
        return self.grid.ravel()[cell] != 1

    def _levy(self, rs, cur, scale):                                            # MPA.py:250-264
        import math
        u = rs.normalvariate(0, self.sigma)
        v = rs.normalvariate(0, 1)
        if abs(v) < 1e-9:
            v = 1e-9
        step = 0.05 * u / (abs(v) ** (1 / self.beta)) * scale
        lim = max(self.R, self.C) * 0.5
        step = min(max(step, -lim), lim)
        angle = rs.uniform(0, 2 * math.pi)
        dr, dc = int(round(step * math.sin(angle))), int(round(step * math.cos(angle)))
        r, c = divmod(cur, self.C)
        return max(0, min(self.R - 1, r + dr)) * self.C + max(0, min(self.C - 1, c + dc))

    def _brownian(self, rs, cur, elite_node, scale):                            # MPA.py:266-282
        import math
        r, c = divmod(cur, self.C)
        if rs.random() < 0.7 and elite_node is not None:
            er, ec = divmod(elite_node, self.C)
            dr, dc = er - r, ec - c
            dist = math.sqrt(dr ** 2 + dc ** 2)
            if dist > 1e-6:
                b = abs(rs.normalvariate(0, 1))
                ms = min(dist, max(1, int(round(scale * b * 5))))
                tr_, tc_ = r + int(round(dr / dist * ms)), c + int(round(dc / dist * ms))
            else:
                return elite_node
        else:
            m = max(1, int(round(max(self.R, self.C) * 0.1 * scale * abs(rs.normalvariate(0, 1)))))
            tr_, tc_ = r + rs.randint(-m, m), c + rs.randint(-m, m)
        return max(0, min(self.R - 1, tr_)) * self.C + max(0, min(self.C - 1, tc_))

    def _reconstruct(self, rs, P, E, idx, levy, scale):                         # MPA.py:284-318
        cur = P[idx]
        prefix = P[:idx + 1]
        avoid = set(prefix[:-1])
        if levy:
            inter = self._levy(rs, cur, scale)
        else:
            node = E[rs.below(len(E))] if E else None
            inter = self._brownian(rs, cur, node, scale)
        full = list(prefix)
        a_start = cur
        if self._free(inter) and inter != a_start:
            seg = self._a_star(a_start, inter, avoid)
            if len(seg) > 1:
                full.extend(seg[1:]); a_start = inter; avoid |= set(seg[1:])
        if a_start != self.T:
            seg = self._a_star(a_start, self.T, avoid)
            if len(seg) > 1:
                full.extend(seg[1:])
        if not full or full[0] != self.S or full[-1] != self.T:
            return self._ind(P)
        return self._ind(full)

    def solve(self):
        p0 = self._a_star(self.S, self.T)                                       # MPA.py:231-245
        if not p0:
            p0 = [self.S, self.T] if self._free(self.T) else [self.S]
        pop = [self._ind(p0) for _ in range(self.N)]
        fit = lambda x: x["stats"][4]
        pop.sort(key=fit)
        best = dict(pop[0])
        self.curve.append(fit(best))
        N, K = self.N, self.K
        for it in range(1, K + 1):                                              # MPA.py:332-440
            pop.sort(key=fit)
            elite = pop[0]
            ratio = it / K
            CF = 0.0 if ratio >= 1.0 else ((1.0 - ratio) ** (2.0 * ratio) if ratio > 0 else 1.0)
            new = []
            for i in range(N):
                rs = _Stream(self.seed, 9, it, i)
                if it <= K / 3:
                    P, E, levy, scale, gate = pop[i], elite, False, self.P, self.P
                    if len(P["path"]) <= 1:
                        new.append(P); continue
                    keep = P
                elif it <= 2 * K / 3:
                    levy = i < N // 2
                    P, E = (pop[i], elite) if levy else (elite, pop[i])
                    scale = gate = self.P if levy else self.P * CF
                    keep = P
                else:
                    P, E, levy, scale, gate, keep = elite, pop[i], True, self.P * CF, self.P * CF, elite
                if len(P["path"]) <= 1:
                    new.append(keep); continue
                idx = rs.randint(0, len(P["path"]) - 2)
                if rs.random() < gate:
                    new.append(self._reconstruct(rs, P["path"], E["path"], idx, levy, scale))
                else:
                    new.append(keep)
            temp = [new[i] if fit(new[i]) < fit(pop[i]) else pop[i] for i in range(N)]          # :380-384
            pop = []
            for i, ind in enumerate(temp):                                      # FADs :387-410
                rs = _Stream(self.seed, 10, it, i)
                final = ind
                if rs.random() < self.FADs:
                    if rs.random() < CF:
                        node = rs.randint(0, self.R - 1) * self.C + rs.randint(0, self.C - 1)
                        if self._free(node):
                            p1 = self._a_star(self.S, node)
                            if p1:
                                p2 = self._a_star(node, self.T, set(p1[:-1]))
                                if p2:
                                    raw = p1 + p2[1:]
                                    if raw and raw[-1] == self.T:
                                        cand = self._ind(raw)
                                        if fit(cand) < fit(final):
                                            final = cand
                    else:
                        pr = self._a_star(self.S, self.T)
                        if pr:
                            cand = self._ind(pr)
                            if fit(cand) < fit(final):
                                final = cand
                pop.append(final)
            pop.sort(key=fit)
            cur = pop[0]
            eq = lambda a, b: abs(a - b) < 1e-9
            cs, bs = cur["stats"], best["stats"]
            if cs[4] < bs[4]:
                best = cur
            elif eq(cs[4], bs[4]):                                              # cascade :422-437
                if cs[0] < bs[0] or (eq(cs[0], bs[0]) and cs[1] < bs[1]) or \
                        (eq(cs[0], bs[0]) and eq(cs[1], bs[1]) and cs[2] < bs[2]) or \
                        (eq(cs[0], bs[0]) and eq(cs[1], bs[1]) and eq(cs[2], bs[2]) and cs[3] < bs[3]):
                    best = cur
            self.curve.append(fit(best))
        self.pop = pop
        return best["path"], best["stats"]
