"""Property tests of the oracle's connectors (hypothesis, CPU): whatever the map, a returned path is a
valid 8-connected obstacle-free chain that honours the corner rule and the avoid set; A* (astar.py), the MPA
variant and Dijkstra agree on the optimal cost; statistics are consistent with the path."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

import pyoracle as O


@st.composite
def cases(draw):
    r = draw(st.integers(3, 18))
    c = draw(st.integers(3, 18))
    seed = draw(st.integers(0, 2**31 - 1))
    dens = draw(st.sampled_from([0.0, 0.1, 0.2, 0.3, 0.4]))
    rng = np.random.default_rng(seed)
    g = (rng.random((r, c)) < dens).astype(int)
    free = np.flatnonzero(g.ravel() == 0)
    if len(free) < 2:
        g[:] = 0
        free = np.arange(r * c)
    s, t = rng.choice(free, 2, replace=False)
    g.ravel()[s], g.ravel()[t] = 2, 3
    avoid = rng.choice(r * c, size=int(rng.integers(0, 6)), replace=False)
    return g, int(s), int(t), avoid, draw(st.booleans()), draw(st.booleans())


def _check_path(g, p, s, t, avoid, allow_diag, restrict, avoid_applies_to_ends):
    C = g.shape[1]
    assert p[0] == s and p[-1] == t and len(set(p.tolist())) == len(p)
    flat = g.ravel()
    assert not np.any(flat[p] == 1)
    dr, dc = np.diff(p // C), np.diff(p % C)
    assert np.all(np.abs(dr) <= 1) and np.all(np.abs(dc) <= 1) and np.all((dr != 0) | (dc != 0))
    diag = (dr != 0) & (dc != 0)
    if not allow_diag:
        assert not diag.any()
    if restrict and diag.any():
        r0, c0 = p[:-1] // C, p[:-1] % C
        assert not np.any(diag & ((g[r0 + dr, c0] == 1) | (g[r0, c0 + dc] == 1)))
    inner = set(p[1:-1].tolist()) if not avoid_applies_to_ends else set(p[1:].tolist())
    assert not (inner & set(int(a) for a in avoid if a not in (s, t) or avoid_applies_to_ends))
    return float(np.where(diag, math.sqrt(2.0), 1.0).sum())


@settings(max_examples=150, deadline=None, derandomize=True)
@given(cases())
def test_connector_paths_are_valid_and_costs_agree(case):
    g, s, t, avoid, allow_diag, restrict = case
    orc = O.AStarOracle(g, allow_diag, restrict)
    bits = O.cells_to_bits(avoid, g.size)
    p0, g0, _, _ = orc.solve(0, s, t, bits)
    p2, g2, _, _ = orc.solve(2, s, t, bits)
    assert (len(p0) == 0) == (len(p2) == 0)                     # same reachability for A* and Dijkstra
    if len(p0):
        c0 = _check_path(g, p0, s, t, avoid, allow_diag, restrict, False)
        c2 = _check_path(g, p2, s, t, avoid, allow_diag, restrict, False)
        assert abs(c0 - g0) < 1e-9 and abs(c2 - g2) < 1e-9 and abs(g0 - g2) < 1e-9     # both optimal
        stats = O.path_stats(g, p0, 0.3, 0.8, 1.8, 100.0)
        assert abs(stats[0] - c0) < 1e-9 and stats[4] >= stats[0] - 1e-12
        if restrict:
            assert stats[3] == 0.0                              # a corner-respecting connector is never charged
    # MPA variant: the avoid set filters every neighbour (target included), no start/target exemption
    p1, g1, _, _ = orc.solve(1, s, t, bits)
    if len(p1):
        _check_path(g, p1, s, t, avoid, allow_diag, restrict, True)
        if len(p0) and t not in set(int(a) for a in avoid):
            assert g1 >= g0 - 1e-9                              # never better than the optimum


@settings(max_examples=60, deadline=None, derandomize=True)
@given(cases(), st.integers(1, 4))
def test_waypoint_chain_properties(case, W):
    g, s, t, _, _, _ = case
    rng = np.random.default_rng(g.size + W)
    free = np.flatnonzero(g.ravel() != 1)
    wps = free[rng.integers(0, len(free), (3, W))].astype(np.int32)
    # a chain may pass a cell again in a later segment (e.g. a waypoint equal to the start): up to 2*R*C cells
    cells, ncell, stats, _ = O.waypoint_fitness(g, wps, 0.3, 0.8, 1.8, 100.0, max_cells=2 * g.size)
    for i in range(3):
        if ncell[i] == 0:
            assert np.isinf(stats[i, 0]) and np.isinf(stats[i, 4])
            continue
        p = cells[i, :ncell[i]]
        assert p[0] == s and p[-1] == t
        assert not np.any(g.ravel()[p] == 1)
        pos = {int(c): k for k, c in enumerate(p.tolist())}
        order = [pos.get(int(w), -1) for w in wps[i]]
        assert all(o >= 0 for o in order)                       # every waypoint is on the path ...
        assert stats[i, 4] >= stats[i, 0]                       # ... and penalties only add


# ---- the ranking shortcut of the MAACO tour kernels (mpp_maaco_rank) -------------------
def _literal_pool(attr, greedy):
    """Candidate indices the final uniform draw picks from, by the literal rules MAACO.py:241-254
    (attr: attractiveness of the candidates in move order)."""
    if greedy:                                             # :241-250
        mx, best = -1.0, []
        for i, a in enumerate(attr):
            if a > mx:
                mx, best = a, [i]
            elif abs(a - mx) < 1e-9:
                best.append(i)
        return best
    assert sum(attr) < 1e-9                                # :252-254: the roulette degenerates to uniform
    return list(range(len(attr)))


@settings(max_examples=300, deadline=None, derandomize=True)
@given(st.lists(st.floats(min_value=0.0, max_value=9.9e-11, allow_nan=False), min_size=8, max_size=8),
       st.integers(1, 255), st.booleans(), st.sampled_from([0, 1, 2]))
def test_rank_order_decides_selection_when_attractiveness_is_tiny(vals, cand_mask, greedy, dup):
    """With every attractiveness < 1e-10 the pool of MAACO.py:241-254 is a function of the ORDER of the
    values only: greedy = the best-ranked candidate (ties -> lower move index) and every later candidate."""
    vals = list(vals)
    if dup:                                                # force ties
        vals[dup] = vals[0]
    cand = [m for m in range(8) if (cand_mask >> m) & 1]
    attr = [vals[m] for m in cand]
    want = [cand[i] for i in _literal_pool(attr, greedy)]
    # ranking as mpp_maaco_rank_kernel builds it: position = number of moves sorting before (value desc, index asc)
    pos = [sum((vals[q] > vals[m]) or (vals[q] == vals[m] and q < m) for q in range(8)) for m in range(8)]
    if greedy:
        best = min(cand, key=lambda m: pos[m])
        got = [m for m in cand if m >= best]
    else:
        got = cand
    assert got == want
