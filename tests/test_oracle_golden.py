"""The C oracle (oracle/mpp_oracle.c) against golden vectors produced by the unmodified reference
under the injected Philox streams (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import pyoracle as O
from conftest import MAACO_CASES, MAACO_DEFAULT_CASES, load_golden

KAT = [  # Random123 known-answer vectors for Philox4x32-10
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        assert O.philox(ctr, key) == want


def test_uniform_mapping_in_unit_interval():
    us = [O.stream_uniform(7, 1, it, ind, d) for it in range(3) for ind in range(5) for d in range(8)]
    assert all(0.0 <= u < 1.0 for u in us) and len(set(us)) == len(us)


@pytest.mark.parametrize("name", MAACO_CASES + MAACO_DEFAULT_CASES)
def test_maaco_oracle_reproduces_reference(name):
    g = load_golden("maaco_" + name)
    N, K = int(g["N"]), int(g["K"])
    o = O.MaacoOracle(g["grid"].astype(int), N, K, seed=int(g["seed"]), **g["params"])
    assert np.array_equal(o.tau.reshape(g["tau0"].shape), g["tau0"])
    pos = 0
    for it in range(1, K + 1):
        cells, ncell, length, turns, _ = o.iterate(it)
        assert np.array_equal(ncell, g["n_cells"][it - 1])
        assert np.array_equal(length, g["length"][it - 1])          # bit-exact fp64 (inf == inf)
        assert np.array_equal(turns, g["turns"][it - 1])
        for a in range(N):
            n = ncell[a]
            assert np.array_equal(cells[a, :n], g["cells"][pos:pos + n])
            pos += n
        assert np.array_equal(o.tau.reshape(g["tau0"].shape), g["tau"][it - 1])
    assert np.array_equal(o.best_path, g["best_cells"])
    assert o.best_len == float(g["best_len"]) and o.best_turns == int(g["best_turns"])
    curve = np.array([np.inf if v is None else v for v in o.curve])
    assert np.array_equal(curve, g["curve"])


def _astar_golden():
    g = load_golden("astar_cases")
    pos = 0
    for i in range(len(g["src"])):
        n, m = g["shape"][i]
        grid = g["grid"][pos:pos + n * m].reshape(n, m).astype(int)
        pos += n * m
        avoid = g["avoid"][g["avoid_off"][i]:g["avoid_off"][i + 1]]
        yield (i, grid, int(g["src"][i]), int(g["dst"][i]), avoid, bool(g["flags"][i][0]), bool(g["flags"][i][1]),
               g["path0"][g["off0"][i]:g["off0"][i + 1]], g["path1"][g["off1"][i]:g["off1"][i + 1]], float(g["g1"][i]))


def test_astar_oracle_reproduces_reference():
    n = 0
    for i, grid, src, dst, avoid, ad, rs, p0, p1, g1 in _astar_golden():
        orc = O.AStarOracle(grid, ad, rs)
        bits = O.cells_to_bits(avoid, grid.size)
        q0, _, _, _ = orc.solve(0, src, dst, bits)
        q1, og1, _, _ = orc.solve(1, src, dst, bits)
        assert np.array_equal(q0, p0), f"case {i} astar.py variant"
        assert np.array_equal(q1, p1) and og1 == g1, f"case {i} MPA variant"
        n += 1
    assert n >= 200


FIT_KEYS = [f"{m}_msd{d}" for m in ("fig7", "blocks40", "blocks64", "rect") for d in (1.8, 2.5)]


@pytest.mark.parametrize("key", FIT_KEYS)
def test_fitness_oracle_reproduces_reference(key):
    g = load_golden("fitness_cases")
    grid = g[key + "_grid"].astype(int)
    msd = float(key.split("msd")[1])
    cells, ncell, stats, _ = O.waypoint_fitness(grid, g[key + "_wps"], 0.3, 0.8, msd, 100.0)
    offs = g[key + "_offs"]
    for i in range(len(ncell)):
        want = g[key + "_paths"][offs[i]:offs[i + 1]]
        assert np.array_equal(cells[i, :ncell[i]], want)
        assert np.array_equal(stats[i], g[key + "_stats"][i])              # bit-exact incl. inf
        ms = O.path_stats(grid, want, 0.1, 0.8, msd, 100.0, True, mode=1)
        assert np.array_equal(ms, g[key + "_mpastats"][i])


def _dijkstra_golden():
    g = load_golden("dijkstra_cases")
    pos = 0
    for i in range(len(g["src"])):
        n, m = g["shape"][i]
        grid = g["grid"][pos:pos + n * m].reshape(n, m).astype(int)
        pos += n * m
        avoid = g["avoid"][g["avoid_off"][i]:g["avoid_off"][i + 1]]
        yield (i, grid, int(g["src"][i]), int(g["dst"][i]), avoid, bool(g["flags"][i][0]), bool(g["flags"][i][1]),
               g["path"][g["off"][i]:g["off"][i + 1]])


def test_dijkstra_oracle_reproduces_reference():
    for i, grid, src, dst, avoid, ad, rs, want in _dijkstra_golden():
        orc = O.AStarOracle(grid, ad, rs)
        got, _, _, _ = orc.solve(2, src, dst, O.cells_to_bits(avoid, grid.size))
        assert np.array_equal(got, want), f"case {i}"


SOLVER_CASES = [(m, n) for m in ("fig7", "blocks40") for n in (20, 33)]


@pytest.mark.parametrize("name,N", SOLVER_CASES)
def test_pso_oracle_reproduces_reference(name, N):
    """Oracle mirror of the PSO solve loop (sequential, asynchronous gbest) vs the reference trajectory."""
    from py_solvers import PsoOracle
    g = load_golden("solver_cases")
    k = f"pso_{name}_{N}"
    _, K, seed = (int(x) for x in g[k + "_meta"])
    o = PsoOracle(g[k + "_grid"].astype(int), K, N, 5, 0.7, 1.5, 1.5, 0.3, 0.8, 1.8, 100.0, seed)
    path, fit = o.solve()
    assert np.array_equal(np.array(o.curve), g[k + "_curve"])
    assert np.array_equal(path, g[k + "_best"]) and fit == g[k + "_stats"][4]
    assert np.array_equal(o.pos, g[k + "_pos"]) and np.array_equal(o.vel, g[k + "_vel"])
    assert np.array_equal(o.pbest_fit, g[k + "_pbest_fit"]) and np.array_equal(o.cur_fit, g[k + "_cur_fit"])


@pytest.mark.parametrize("name,N", SOLVER_CASES)
def test_ga_oracle_reproduces_reference(name, N):
    from py_solvers import GaOracle
    g = load_golden("solver_cases")
    k = f"ga_{name}_{N}"
    _, K, seed = (int(x) for x in g[k + "_meta"])
    o = GaOracle(g[k + "_grid"].astype(int), K, N, 5, 0.1, 0.8, 3, 0.3, 0.8, 1.8, 100.0, seed)
    path, st = o.solve()
    assert np.array_equal(np.array(o.curve), g[k + "_curve"])
    assert np.array_equal(path, g[k + "_best"]) and np.array_equal(st, g[k + "_stats"])
    assert np.array_equal(np.array([ind[0] for ind in o.pop], np.int32), g[k + "_chrom"])
    assert np.array_equal(np.array([ind[1][4] for ind in o.pop]), g[k + "_fit"])


@pytest.mark.parametrize("name,N", [(m, n) for m in ("fig7", "blocks40") for n in (20, 24)])
def test_mpa_oracle_reproduces_reference(name, N):
    """Oracle mirror of the MPA solve loop (Levy/Brownian targets with libm, private A*, memory, FADs, sorts)."""
    from py_solvers import MpaOracle
    g = load_golden("solver_cases")
    k = f"mpa_{name}_{N}"
    _, K, seed, beta10 = (int(x) for x in g[k + "_meta"])
    o = MpaOracle(g[k + "_grid"].astype(int), N, K, 0.2, 0.5, beta10 / 10.0, 0.1, 0.8, 1.8, 100.0, seed)
    path, st = o.solve()
    assert np.array_equal(np.array(o.curve), g[k + "_curve"])
    assert np.array_equal(np.array(path, np.int32), g[k + "_best"]) and np.array_equal(st, g[k + "_stats"])
    assert np.array_equal(np.array([ind["stats"][4] for ind in o.pop]), g[k + "_pop_fit"])
    offs = g[k + "_pop_offs"]
    for i, ind in enumerate(o.pop):
        assert np.array_equal(np.array(ind["path"], np.int32), g[k + "_pop_cells"][offs[i]:offs[i + 1]])
