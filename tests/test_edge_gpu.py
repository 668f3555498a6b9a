"""Edge cases the reference's semantics cover: degenerate shapes (1xN, Nx1), adjacent / unreachable endpoints,
4-connected and corner-cutting policies, tiny colonies, ragged batches.  CUDA vs oracle, bit-exact."""
import numpy as np
import pytest

from conftest import MAACO_DEFAULT

pytestmark = pytest.mark.gpu


def _grid(shape, s, t, obstacles=()):
    g = np.zeros(shape, int)
    for o in obstacles:
        g[o] = 1
    g[s], g[t] = 2, 3
    return g


GRIDS = {
    "row": _grid((1, 17), (0, 0), (0, 16)),
    "col": _grid((23, 1), (22, 0), (0, 0)),
    "adjacent": _grid((4, 4), (1, 1), (2, 2)),
    "blocked": _grid((6, 7), (0, 0), (5, 6), [(r, 3) for r in range(6)]),
    "reverse": _grid((15, 33), (14, 32), (0, 0), [(7, c) for c in range(4, 33)]),           # T up-left of S
    "snake": _grid((9, 9), (0, 0), (8, 8), [(2, c) for c in range(0, 8)] + [(5, c) for c in range(1, 9)]),
    "wide33": _grid((5, 33), (2, 0), (2, 32), [(1, 10), (2, 10), (3, 10)]),                 # pitch crosses a word
}


@pytest.mark.parametrize("name", list(GRIDS))
def test_maaco_edge_grids(name):
    import pyoracle as O
    from maaco_path_planing_b200 import MAACO
    g = GRIDS[name]
    N, K = 13, 4                                                          # colony size not a multiple of 4 / 32
    dev = MAACO(g, N, K, rng_seed=5, verbose=False, **MAACO_DEFAULT)
    orc = O.MaacoOracle(g, N, K, seed=5, **MAACO_DEFAULT)
    for it in range(1, K + 1):
        dev.run_iteration(it)
        nc, ln, tn, cells = dev.last_tours()
        ocells, onc, oln, otn, _ = orc.iterate(it)
        assert np.array_equal(nc, onc) and np.array_equal(ln, oln) and np.array_equal(tn, otn)
        for a in range(N):
            assert np.array_equal(cells[a, :nc[a]], ocells[a, :onc[a]])
        assert np.array_equal(dev.pheromone_matrix.ravel(), orc.tau)
    d2 = MAACO(g, N, K, rng_seed=5, verbose=False, **MAACO_DEFAULT)
    path, length, turns = d2.solve_path_planning()
    o2 = O.MaacoOracle(g, N, K, seed=5, **MAACO_DEFAULT)
    op, ol, ot = o2.solve()
    C = g.shape[1]
    assert [r * C + c for r, c in path] == list(op) and length == ol
    assert turns == (ot if ot >= 0 else float("inf"))


@pytest.mark.parametrize("name", list(GRIDS))
@pytest.mark.parametrize("allow_diag,restrict", [(True, True), (True, False), (False, True)])
def test_astar_edge_grids_and_policies(name, allow_diag, restrict):
    import pyoracle as O
    from maaco_path_planing_b200 import GridMap
    from maaco_path_planing_b200.engine import SearchEngine
    g = GRIDS[name]
    R, C = g.shape
    eng = SearchEngine(GridMap(g))
    orc = O.AStarOracle(g, allow_diag, restrict)
    rng = np.random.default_rng(R * 100 + C)
    n = 40
    src = rng.integers(0, R * C, n).astype(np.int32)
    dst = rng.integers(0, R * C, n).astype(np.int32)
    s0, t0 = int(np.flatnonzero(g.ravel() == 2)[0]), int(np.flatnonzero(g.ravel() == 3)[0])
    src[0], dst[0] = s0, t0
    src[1], dst[1] = t0, s0
    src[2], dst[2] = s0, s0
    avoid = np.zeros((n, (R * C + 31) // 32), np.uint32)
    for i in range(3, n):
        for c in rng.integers(0, R * C, int(rng.integers(0, 6))):
            avoid[i, c >> 5] |= np.uint32(1 << (c & 31))
    for variant in (0, 1):
        cells, ncell, gg = eng.astar_batch(variant, src, dst, avoid.view(np.int32), allow_diag, restrict)
        cells, ncell, gg = cells.cpu().numpy(), ncell.cpu().numpy(), gg.cpu().numpy()
        for i in range(n):
            want, wg, _, _ = orc.solve(variant, int(src[i]), int(dst[i]), avoid[i])
            assert ncell[i] == len(want) and np.array_equal(cells[i, :ncell[i]], want), (variant, i)
            assert gg[i] == wg or (np.isinf(gg[i]) and np.isinf(wg))


def test_fitness_zero_waypoints_and_single_individual():
    import pyoracle as O
    from maaco_path_planing_b200 import GridMap, blocks_map
    from maaco_path_planing_b200.engine import SearchEngine, make_policy
    g = blocks_map(40, 0.2, seed=8)
    eng = SearchEngine(GridMap(g))
    pol = make_policy(0.3, 0.8, 1.8, 100.0)
    wps = np.zeros((1, 0), np.int32)                                      # W = 0: direct start -> target connector
    cells, ncell, stats = eng.waypoint_fitness(wps, pol)
    ocells, oncell, ostats, _ = O.waypoint_fitness(g, wps, 0.3, 0.8, 1.8, 100.0)
    n = int(ncell[0])
    assert n == oncell[0] and np.array_equal(cells[0, :n].cpu().numpy(), ocells[0, :n])
    assert np.array_equal(stats.cpu().numpy(), ostats)


def test_small_path_buffers_grow():
    """max_cells / heap smaller than needed: the engine enlarges and repeats instead of truncating silently."""
    import pyoracle as O
    from maaco_path_planing_b200 import GridMap, blocks_map
    from maaco_path_planing_b200.engine import SearchEngine, make_policy
    g = blocks_map(64, 0.2, seed=9)
    eng = SearchEngine(GridMap(g), max_cells=16, heap_cap=64)
    rng = np.random.default_rng(0)
    free = np.flatnonzero(g.ravel() != 1)
    wps = free[rng.integers(0, len(free), (32, 4))].astype(np.int32)
    cells, ncell, stats = eng.waypoint_fitness(wps, make_policy(0.3, 0.8, 1.8, 100.0))
    _, oncell, ostats, _ = O.waypoint_fitness(g, wps, 0.3, 0.8, 1.8, 100.0)
    assert np.array_equal(ncell.cpu().numpy(), oncell) and np.array_equal(stats.cpu().numpy(), ostats)
    assert eng.max_cells > 16
