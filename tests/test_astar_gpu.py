"""CUDA connectors / waypoint-chain fitness / path statistics (through the C ABI) vs the reference's
golden vectors and the oracle.  Paths bit-exact; g, length, penalties, fitness bit-exact fp64
(north-star tolerance 1e-9 relative is therefore met with margin)."""
import numpy as np
import pytest

from conftest import load_golden
from test_oracle_golden import FIT_KEYS, _astar_golden

pytestmark = pytest.mark.gpu


def _engine(grid, **kw):
    from maaco_path_planing_b200 import GridMap
    from maaco_path_planing_b200.engine import SearchEngine
    return SearchEngine(GridMap(grid), **kw)


def test_astar_batch_matches_reference_goldens():
    import pyoracle as O
    for i, grid, src, dst, avoid, ad, rs, p0, p1, g1 in _astar_golden():
        eng = _engine(grid)
        bits = O.cells_to_bits(avoid, grid.size).view(np.int32)[None, :]
        for variant, want in ((0, p0), (1, p1)):
            cells, ncell, g = eng.astar_batch(variant, [src], [dst], bits, ad, rs)
            n = int(ncell[0])
            assert n == len(want), f"case {i} variant {variant}: {n} vs {len(want)}"
            assert np.array_equal(cells[0, :n].cpu().numpy(), want), f"case {i} variant {variant}"
            if variant == 1:
                assert float(g[0]) == g1


@pytest.mark.parametrize("size,dens,seed", [(48, 0.25, 1), (100, 0.2, 2), (37, 0.3, 3)])
def test_astar_batch_random_vs_oracle(size, dens, seed):
    """Many searches on one map in one launch (both variants, with per-search avoid sets)."""
    import pyoracle as O
    rng = np.random.default_rng(seed)
    grid = (rng.random((size, size + 5)) < dens).astype(int)
    n = 96
    R, C = grid.shape
    src = rng.integers(0, R * C, n).astype(np.int32)
    dst = rng.integers(0, R * C, n).astype(np.int32)
    dst[:4] = src[:4]
    avoid = np.zeros((n, (R * C + 31) // 32), np.uint32)
    for i in range(n):
        for c in rng.integers(0, R * C, int(rng.integers(0, 60))):
            avoid[i, c >> 5] |= np.uint32(1 << (c & 31))
    eng = _engine(grid)
    orc = O.AStarOracle(grid, True, True)
    for variant in (0, 1):
        cells, ncell, g = eng.astar_batch(variant, src, dst, avoid.view(np.int32), True, True)
        cells, ncell, g = cells.cpu().numpy(), ncell.cpu().numpy(), g.cpu().numpy()
        for i in range(n):
            want, wg, _, _ = orc.solve(variant, int(src[i]), int(dst[i]), avoid[i])
            assert ncell[i] == len(want), (variant, i)
            assert np.array_equal(cells[i, :ncell[i]], want), (variant, i)
            assert g[i] == wg or (np.isinf(g[i]) and np.isinf(wg)), (variant, i, g[i], wg)


@pytest.mark.parametrize("key", FIT_KEYS)
def test_waypoint_fitness_matches_reference_goldens(key):
    from maaco_path_planing_b200.engine import make_policy
    g = load_golden("fitness_cases")
    grid = g[key + "_grid"].astype(int)
    msd = float(key.split("msd")[1])
    eng = _engine(grid)
    pol = make_policy(0.3, 0.8, msd, 100.0)
    cells, ncell, stats = eng.waypoint_fitness(g[key + "_wps"], pol)
    cells, ncell, stats = cells.cpu().numpy(), ncell.cpu().numpy(), stats.cpu().numpy()
    offs = g[key + "_offs"]
    for i in range(len(ncell)):
        want = g[key + "_paths"][offs[i]:offs[i + 1]]
        assert ncell[i] == len(want)
        assert np.array_equal(cells[i, :ncell[i]], want)
        assert np.array_equal(stats[i], g[key + "_stats"][i]), (i, stats[i], g[key + "_stats"][i])
    # MPA-mode statistics of the same paths (MPA.py:215-229)
    P = len(ncell)
    buf = np.zeros((P, max(1, int(ncell.max()))), np.int32)
    for i in range(P):
        buf[i, :ncell[i]] = g[key + "_paths"][offs[i]:offs[i + 1]]
    ms = eng.path_stats(buf, ncell, make_policy(0.1, 0.8, msd, 100.0, mode=1)).cpu().numpy()
    assert np.array_equal(ms, g[key + "_mpastats"])


@pytest.mark.parametrize("size,N,W", [(100, 256, 5), (256, 128, 5), (64, 200, 3), (512, 192, 5)])
def test_waypoint_fitness_vs_oracle(size, N, W):
    """Population-sized batches incl. obstacle / duplicate waypoints (invalid individuals)."""
    import pyoracle as O
    from maaco_path_planing_b200 import blocks_map
    from maaco_path_planing_b200.engine import make_policy
    grid = blocks_map(size, 0.2, seed=size)
    rng = np.random.default_rng(size)
    wps = rng.integers(0, size * size, (N, W)).astype(np.int32)        # PSO-style: may hit obstacles
    free = np.flatnonzero(grid.ravel() != 1)
    wps[N // 4:] = free[rng.integers(0, len(free), (N - N // 4, W))]    # GA-style: free cells
    wps[5, 1] = wps[5, 0]
    eng = _engine(grid)
    cells, ncell, stats = eng.waypoint_fitness(wps, make_policy(0.3, 0.8, 1.8, 100.0))
    cells, ncell, stats = cells.cpu().numpy(), ncell.cpu().numpy(), stats.cpu().numpy()
    ocells, oncell, ostats, oexp = O.waypoint_fitness(grid, wps, 0.3, 0.8, 1.8, 100.0, threads=0)
    assert np.array_equal(ncell, oncell)
    for i in range(N):
        assert np.array_equal(cells[i, :ncell[i]], ocells[i, :oncell[i]]), i
    assert np.array_equal(stats, ostats)
    assert (ncell > 0).sum() > N // 3
    assert eng.expansions()[0] == oexp                                  # same number of node expansions


def test_path_stats_edge_cases():
    import pyoracle as O
    from maaco_path_planing_b200.engine import make_policy
    grid = np.zeros((9, 11), int)
    grid[4, 3:8] = 1
    grid[0, 0], grid[8, 10] = 2, 3
    eng = _engine(grid)
    C = 11
    paths = [[], [5], [0, 8 * C + 10], [0, 1, 2, C + 3, 2 * C + 3, 3 * C + 2, 3 * C + 3, 3 * C + 4],
             [3 * C + 2, 4 * C + 2 - C + 1], list(range(0, 11)), [3 * C + 8, 4 * C + 8 + 0, 5 * C + 7]]
    buf = np.zeros((len(paths), 16), np.int32)
    n = np.array([len(p) for p in paths], np.int32)
    for i, p in enumerate(paths):
        buf[i, :len(p)] = p
    for mode, tpf in ((0, 0.3), (1, 0.1)):
        for msd in (1.8, 3.2, 0.0, 16.0, 23.7):                     # (>= 15.9 was rejected in round 1)
            st = eng.path_stats(buf, n, make_policy(tpf, 0.8, msd, 100.0, mode=mode)).cpu().numpy()
            for i, p in enumerate(paths):
                want = O.path_stats(grid, p, tpf, 0.8, msd, 100.0, True, mode=mode)
                assert np.array_equal(st[i], want), (mode, msd, i, st[i], want)


def test_dijkstra_matches_reference_goldens_and_anchor():
    """dijkstra.py:32-97 = connector variant 2 (zero heuristic).  Anchor (main.py policy):
    T=12, SP=0.3274397055203064, fitness=35.41830095052029."""
    import pyoracle as O
    from test_oracle_golden import _dijkstra_golden
    from maaco_path_planing_b200.dijkstra import DijkstraSolver
    for i, grid, src, dst, avoid, ad, rs, want in _dijkstra_golden():
        eng = _engine(grid)
        bits = O.cells_to_bits(avoid, grid.size).view(np.int32)[None, :]
        cells, ncell, g = eng.astar_batch(2, [src], [dst], bits, ad, rs)
        n = int(ncell[0])
        assert n == len(want) and np.array_equal(cells[0, :n].cpu().numpy(), want), f"case {i}"
    grid = load_golden("env_grids")["fig7"].astype(int)
    d = DijkstraSolver(grid, turn_penalty_factor=0.3, safety_penalty_factor=0.8, min_safe_distance=1.8,
                       diagonal_obstacle_penalty_value=100.0)
    path, L, T, SP, DP, F = d.solve()
    assert T == 12 and SP == 0.3274397055203064 and F == 35.41830095052029 and len(path) == 28


def test_config3_full_population_properties():
    """BASELINE config 3 at full size (4096 individuals x 5 waypoints, 512x512): size-independent properties --
    idempotence (same chromosomes -> identical bytes), every valid path is a start->target chain through its
    waypoints in order, statistics recomputed by the stand-alone kernel equal the fused ones."""
    from maaco_path_planing_b200 import blocks_map
    from maaco_path_planing_b200.engine import make_policy
    size, N, W = 512, 4096, 5
    grid = blocks_map(size, 0.2, seed=3000 + size)
    rng = np.random.default_rng(11)
    free = np.flatnonzero(grid.ravel() != 1)
    wps = free[rng.integers(0, len(free), (N, W))].astype(np.int32)
    eng = _engine(grid)
    pol = make_policy(0.3, 0.8, 1.8, 100.0)
    cells, ncell, stats = eng.waypoint_fitness(wps, pol)
    c2, n2, s2 = eng.waypoint_fitness(wps, pol)
    assert bool((ncell == n2).all()) and bool((stats == s2).all() | (stats != stats).all())
    nc = ncell.cpu().numpy()
    st = stats.cpu().numpy()
    assert np.array_equal(st, s2.cpu().numpy())
    assert (nc > 0).mean() > 0.95
    re = eng.path_stats(cells, ncell, pol).cpu().numpy()
    ok = nc > 0
    assert np.array_equal(re[ok], st[ok])                                  # fused stats == stand-alone kernel
    cc = cells.cpu().numpy()
    flat = grid.ravel()
    for i in np.flatnonzero(ok)[:128]:
        p = cc[i, :nc[i]]
        assert p[0] == 0 and p[-1] == size * size - 1 and not np.any(flat[p] == 1)
        pos = {int(c): k for k, c in enumerate(p.tolist())}
        order = [pos[int(w)] for w in wps[i]]
        assert order == sorted(order)                                      # waypoints visited in chromosome order
        assert np.array_equal(c2[i, :nc[i]].cpu().numpy(), p)
