"""Reference harness -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Loads the *unmodified* reference (dvnam1605/MAACO-path-planing) from
``$MAACO_REF_DIR`` or ``/root/reference`` under a matplotlib stub and rebinds each
module's ``random`` name (and ``MAACO.np``) to a counter-based Philox4x32-10
"tape" so that every uniform draw is attributable to a stream
``(seed, class, iteration, individual)`` with an in-stream cursor.  The CUDA
kernels, the C oracle (``oracle/mpp_oracle.c``) and this shim implement the same
stream contract (DESIGN.md "RNG contract"), which is what makes "bit-exact
under injected uniform draws" checkable.

This file only *exists to generate golden vectors* (``tests/golden/make_golden.py``)
and to pin the C oracle against the real reference
(``oracle/validate_against_reference.py``).  It runs only where the reference
tree is present (the authoring container); nothing on the GPU box imports it.

Reference RNG call sites covered (file:line in /root/reference):
  MAACO.py:232,250,254,259,262 ; pso.py:50,51,105,160,186,187,189,190 ;
  ga_solver.py:50,51,130,139,145,147,158 ;
  MPA.py:248,254,255,261,267,271,278,279,343,344,358,359,371,372,389,390,391
"""
from __future__ import annotations

import importlib
import math
import os
import random as _pyrandom
import sys
import types

import numpy as np

# ----------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11) -- pure Python ints.
# ----------------------------------------------------------------------------
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, \
                         ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def u53(a, b):
    """Two 32-bit words -> double in [0,1) (same mapping as MT genrand_res53)."""
    return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0


# stream classes (shared contract)
CLS_MAACO_TOUR = 1
CLS_PSO_INIT = 2
CLS_PSO_PAD = 3
CLS_PSO_UPDATE = 4
CLS_GA_INIT = 5
CLS_GA_PAD = 6
CLS_GA_SELECT = 7
CLS_GA_BREED = 8
CLS_MPA_PHASE = 9
CLS_MPA_FADS = 10


def stream_uniform(seed, cls, it, ind, d):
    """Draw number ``d`` of stream (seed, cls, it, ind)."""
    key = (seed & _MASK, (seed >> 32) & _MASK)
    w = philox4x32_10((d >> 1, ind & _MASK, it & _MASK, cls), key)
    return u53(w[0], w[1]) if (d & 1) == 0 else u53(w[2], w[3])


class TapeRandom(_pyrandom.Random):
    """``random``-module stand-in whose uniforms come from the current stream.

    ``_randbelow(n) = floor(u*n)`` replaces CPython's getrandbits rejection so
    that ``choice``/``randint``/``sample`` consume exactly one uniform per
    (re)try; ``normalvariate``/``uniform`` are CPython's own code running over
    ``self.random``.
    """

    def __init__(self, seed=0, locate=None):
        super().__init__(0)
        self.tape_seed = int(seed)
        self.key = None        # (cls, it, ind)
        self.cursor = 0
        self.locate = locate   # callable -> (cls, it, ind) or None (keep)
        self.log = None        # optional list of (cls,it,ind,d,u)
        self.counts = {}

    # -- stream control ------------------------------------------------------
    def set_stream(self, cls, it, ind):
        k = (cls, it, ind)
        if k != self.key:
            self.key = k
            self.cursor = 0

    # -- the only entropy source --------------------------------------------
    def random(self):
        if self.locate is not None:
            k = self.locate()
            if k is not None and k != self.key:
                self.key = k
                self.cursor = 0
        if self.key is None:
            raise RuntimeError("TapeRandom: draw outside any stream")
        cls, it, ind = self.key
        u = stream_uniform(self.tape_seed, cls, it, ind, self.cursor)
        if self.log is not None:
            self.log.append((cls, it, ind, self.cursor, u))
        self.counts[cls] = self.counts.get(cls, 0) + 1
        self.cursor += 1
        return u

    def _randbelow(self, n):
        j = int(self.random() * n)
        return j if j < n else n - 1

    def getrandbits(self, k):  # pragma: no cover - must never be reached
        raise RuntimeError("TapeRandom.getrandbits must not be used")


class _NpRandomProxy:
    def __init__(self, tape):
        self._tape = tape

    def choice(self, n, p=None):
        # numpy legacy RandomState.choice(n, p=p), replace=True, size=None:
        #   validate p, cdf = cumsum(p); cdf /= cdf[-1];
        #   idx = cdf.searchsorted(random_sample(), side='right')
        p = np.asarray(p, dtype=np.float64)
        atol = math.sqrt(np.finfo(np.float64).eps)
        ps = math.fsum(p.tolist())  # numpy uses a Kahan sum; differs <1ulp, only a 1.5e-8 gate
        if np.isnan(ps):
            raise ValueError("probabilities contain NaN")
        if (p < 0).any():
            raise ValueError("probabilities are not non-negative")
        if abs(ps - 1.0) > atol:
            raise ValueError("probabilities do not sum to 1")
        cdf = p.cumsum()
        cdf /= cdf[-1]
        u = self._tape.random()
        return int(cdf.searchsorted(u, side="right"))


class _NpProxy:
    """Forwards everything to numpy except ``.random.choice``."""

    def __init__(self, tape):
        self.random = _NpRandomProxy(tape)

    def __getattr__(self, name):
        return getattr(np, name)


# ----------------------------------------------------------------------------
# Loading the reference
# ----------------------------------------------------------------------------
def ref_dir():
    """$MAACO_REF_DIR -> /root/reference (the authoring container) -> <repo>/baseline/_ref (the git-ignored copy that
    __graft_entry__.build() makes and that travels to the GPU box)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in (os.environ.get("MAACO_REF_DIR"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if d and os.path.isfile(os.path.join(d, "MAACO.py")):
            return d
    raise FileNotFoundError("reference tree not found ($MAACO_REF_DIR, /root/reference, baseline/_ref)")


def _install_matplotlib_stub():
    if "matplotlib" in sys.modules and not getattr(sys.modules["matplotlib"], "_mpp_stub", False):
        return

    class _Anything:
        def __call__(self, *a, **k):
            return self

        def __getattr__(self, n):
            return self

        def __iter__(self):
            return iter(())

    def _mk(name):
        m = types.ModuleType(name)
        m._mpp_stub = True
        m.__getattr__ = lambda attr: _Anything()
        return m

    mpl = _mk("matplotlib")
    mpl.pyplot = _mk("matplotlib.pyplot")
    mpl.colors = _mk("matplotlib.colors")
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = mpl.pyplot
    sys.modules["matplotlib.colors"] = mpl.colors


_REF = None


def load_reference():
    """Import the reference modules (unmodified). Returns a namespace."""
    global _REF
    if _REF is not None:
        return _REF
    for n in ("MAACO", "MPA", "pso", "ga_solver", "astar", "helper", "env", "visualization", "dijkstra"):
        if n in sys.modules and not getattr(sys.modules[n], "__file__", "").startswith(ref_dir()):
            raise RuntimeError(f"module name {n!r} already taken by {sys.modules[n].__file__}; "
                               "the reference harness must run in its own interpreter")
    _install_matplotlib_stub()
    sys.path.insert(0, ref_dir())
    try:
        ns = types.SimpleNamespace()
        for n in ("env", "helper", "astar", "dijkstra", "MAACO", "MPA", "pso", "ga_solver"):
            setattr(ns, n, importlib.import_module(n))
    finally:
        sys.path.remove(ref_dir())
    _REF = ns
    return ns


class quiet:
    """Silence the reference's progress prints."""

    def __enter__(self):
        self._o = sys.stdout
        sys.stdout = open(os.devnull, "w")

    def __exit__(self, *a):
        sys.stdout.close()
        sys.stdout = self._o


# ----------------------------------------------------------------------------
# MAACO under the tape
# ----------------------------------------------------------------------------
def run_maaco(grid, params, seed, record=True):
    """Run reference MAACO.solve_path_planning under the tape.

    Returns dict with the solver, the returned triple, and (if record) per
    iteration: list of (path, length, turns) per ant and tau after the update.
    """
    ref = load_reference()
    M = ref.MAACO
    tape = TapeRandom(seed)
    old_random, old_np = M.random, M.np
    M.random, M.np = tape, _NpProxy(tape)
    try:
        solver = M.MAACO(np.array(grid), **params)
        tau0 = solver.pheromone_matrix.copy()
        trace = {"tours": [], "tau": [], "best": []}
        orig_construct = solver._construct_ant_solution_maaco
        orig_update = solver._update_pheromone_trails_maaco
        cur = {"it": 0, "row": None}

        def construct(ant_id, it):
            tape.set_stream(CLS_MAACO_TOUR, it, ant_id)
            out = orig_construct(ant_id, it)
            if record:
                if cur["it"] != it:
                    cur["it"] = it
                    cur["row"] = []
                    trace["tours"].append(cur["row"])
                cur["row"].append(([(int(r), int(c)) for r, c in out[0]], float(out[1]), out[2]))
            return out

        def update(paths, best_len):
            orig_update(paths, best_len)
            if record:
                trace["tau"].append(solver.pheromone_matrix.copy())
                trace["best"].append((float(solver.best_path_length_overall),
                                      solver.best_path_turns_overall,
                                      [(int(r), int(c)) for r, c in solver.best_path_overall]))

        solver._construct_ant_solution_maaco = construct
        solver._update_pheromone_trails_maaco = update
        with quiet():
            res = solver.solve_path_planning()
        return {"solver": solver, "result": res, "trace": trace, "tau0": tau0, "draws": dict(tape.counts)}
    finally:
        M.random, M.np = old_random, old_np


# ----------------------------------------------------------------------------
# frame-walking stream locators for the solvers that draw inline
# ----------------------------------------------------------------------------
def _find_frame(name, filename_tail):
    f = sys._getframe(2)
    while f is not None:
        co = f.f_code
        if co.co_name == name and co.co_filename.endswith(filename_tail):
            return f
        f = f.f_back
    return None


def make_pso_locator():
    def locate():
        f = _find_frame("_initialize_particles", "pso.py")
        if f is not None:
            if f.f_lineno >= 159:  # padding loop pso.py:159-160
                return (CLS_PSO_PAD, 0, len(f.f_locals["self"].particles))
            return (CLS_PSO_INIT, 0, f.f_locals["attempts"] - 1)
        f = _find_frame("solve", "pso.py")
        if f is not None and 178 <= f.f_lineno <= 206:
            return (CLS_PSO_UPDATE, f.f_locals["iteration"], f.f_locals["p_idx"])
        raise RuntimeError("PSO draw from unexpected site")
    return locate


def make_ga_locator():
    def locate():
        f = _find_frame("_initialize_population", "ga_solver.py")
        if f is not None:
            if f.f_lineno >= 129:
                return (CLS_GA_PAD, 0, len(f.f_locals["self"].population))
            return (CLS_GA_INIT, 0, f.f_locals["attempts"] - 1)
        f = _find_frame("_selection", "ga_solver.py")
        if f is not None:
            g = _find_frame("solve", "ga_solver.py")
            return (CLS_GA_SELECT, g.f_locals["gen"], len(f.f_locals["selected_parents"]))
        f = _find_frame("solve", "ga_solver.py")
        if f is not None and 186 <= f.f_lineno <= 206:
            return (CLS_GA_BREED, f.f_locals["gen"], f.f_locals["idx"] // 2 - 1)
        raise RuntimeError("GA draw from unexpected site")
    return locate


def make_mpa_locator():
    def locate():
        f = _find_frame("solve_path_planning", "MPA.py")
        if f is None:
            raise RuntimeError("MPA draw from unexpected site")
        ln = f.f_lineno
        it = f.f_locals["iter_num_solve_mpa_main"]
        if 340 <= ln <= 347:
            return (CLS_MPA_PHASE, it, f.f_locals["i_p1_main"])
        if 349 <= ln <= 364:
            return (CLS_MPA_PHASE, it, f.f_locals["i_p2_main"])
        if 366 <= ln <= 377:
            return (CLS_MPA_PHASE, it, f.f_locals["i_p3_main"])
        if 386 <= ln <= 410:
            return (CLS_MPA_FADS, it, len(f.f_locals["self"].population))
        raise RuntimeError(f"MPA draw from unexpected line {ln}")
    return locate


def with_tape(module, tape):
    """Context manager: rebind ``module.random`` to the tape."""
    class _Ctx:
        def __enter__(self_):
            self_.old = module.random
            module.random = tape
            return tape

        def __exit__(self_, *a):
            module.random = self_.old
    return _Ctx()


def blocks_map(n, frac=0.20, seed=0, rows=None, cols=None):
    """Synthetic block-obstacle map; S=(0,0), T=(n-1,n-1)."""
    rows = rows or n
    cols = cols or n
    rng = np.random.default_rng(seed)
    g = np.zeros((rows, cols), dtype=np.int64)
    m = max(2, min(rows, cols) // 12)
    while g.mean() < frac:
        h, w = rng.integers(1, m, 2)
        r = rng.integers(0, rows - h)
        c = rng.integers(0, cols - w)
        g[r:r + h, c:c + w] = 1
    g[:2, :2] = 0
    g[-2:, -2:] = 0
    g[0, 0] = 2
    g[rows - 1, cols - 1] = 3
    return g


# ----------------------------------------------------------------------------
# PSO / GA under the tape
# ----------------------------------------------------------------------------
def run_pso(grid, kwargs, seed):
    """Reference PSOSolver.solve under the tape; returns result, curve and final particle state."""
    ref = load_reference()
    tape = TapeRandom(seed, locate=make_pso_locator())
    with with_tape(ref.pso, tape):
        with quiet():
            s = ref.pso.PSOSolver(grid=np.array(grid), **kwargs)
            init = {}
            orig_init = s._initialize_particles

            def init_hook():
                ok = orig_init()
                init["pos"] = np.array([p["position"] for p in s.particles], dtype=float)
                init["vel"] = np.array([p["velocity"] for p in s.particles], dtype=float)
                init["fit"] = np.array([p["current_fitness"] for p in s.particles], dtype=float)
                init["gbest_fit"] = s.gbest_particle_data["fitness"]
                return ok
            s._initialize_particles = init_hook
            res = s.solve()
    C = np.array(grid).shape[1]
    return {"result": res, "curve": list(s.convergence_curve), "init": init, "draws": dict(tape.counts),
            "pos": np.array([p["position"] for p in s.particles], dtype=float),
            "vel": np.array([p["velocity"] for p in s.particles], dtype=float),
            "pbest_fit": np.array([p["pbest_fitness"] for p in s.particles], dtype=float),
            "cur_fit": np.array([p["current_fitness"] for p in s.particles], dtype=float),
            "best_cells": np.array([int(r) * C + int(c) for r, c in res[0]], np.int32)}


def run_ga(grid, kwargs, seed):
    """Reference GASolver.solve under the tape; returns result, curve, initial and final populations."""
    ref = load_reference()
    tape = TapeRandom(seed, locate=make_ga_locator())
    C = np.array(grid).shape[1]
    with with_tape(ref.ga_solver, tape):
        with quiet():
            s = ref.ga_solver.GASolver(grid=np.array(grid), **kwargs)
            init = {}
            orig_init = s._initialize_population

            def init_hook():
                ok = orig_init()
                init["chrom"] = np.array([[r * C + c for r, c in ind["chromosome"]] for ind in s.population], np.int32)
                init["fit"] = np.array([ind["fitness"] for ind in s.population], dtype=float)
                return ok
            s._initialize_population = init_hook
            res = s.solve()
    return {"result": res, "curve": list(s.convergence_curve), "init": init, "draws": dict(tape.counts),
            "chrom": np.array([[r * C + c for r, c in ind["chromosome"]] for ind in s.population], np.int32),
            "fit": np.array([ind["fitness"] for ind in s.population], dtype=float),
            "best_cells": np.array([int(r) * C + int(c) for r, c in res[0]], np.int32)}


def run_mpa(grid, kwargs, seed):
    """Reference MPA.solve_path_planning under the tape; returns result, curve, final population."""
    ref = load_reference()
    tape = TapeRandom(seed, locate=make_mpa_locator())
    C = np.array(grid).shape[1]
    with with_tape(ref.MPA, tape):
        with quiet():
            s = ref.MPA.MPA(grid=np.array(grid), **kwargs)
            res = s.solve_path_planning()
    flat, offs = [], [0]
    for ind in s.population:
        flat.extend(int(r) * C + int(c) for r, c in ind["path"])
        offs.append(len(flat))
    return {"result": res, "curve": [np.inf if v is None else v for v in s.convergence_curve_data],
            "draws": dict(tape.counts), "pop_cells": np.array(flat, np.int32), "pop_offs": np.array(offs, np.int64),
            "pop_fit": np.array([ind["fitness"] for ind in s.population], dtype=float),
            "best_cells": np.array([int(r) * C + int(c) for r, c in res[0]], np.int32)}
