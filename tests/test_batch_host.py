"""Host-side helpers of the batched solvers (no GPU needed)."""
import numpy as np

from maaco_path_planing_b200.batch import _grids_u8, shard_maps


def _old(grids):
    g = np.ascontiguousarray(np.asarray(grids, dtype=int))
    return np.ascontiguousarray(np.clip(g, 0, 255), dtype=np.uint8)


def test_grids_u8_equals_clip_of_int_grid():
    rng = np.random.default_rng(3)
    g = rng.integers(0, 4, (5, 17, 23))                       # int64 cell codes (what blocks_map returns)
    g[0, 1, 1], g[1, 2, 2], g[2, 3, 3] = 300, -7, 255         # out-of-range values clip like np.clip
    for arr in (g, g.astype(np.int32), g.astype(np.int16), np.clip(g, 0, 255).astype(np.uint8), g.astype(float) + 0.25):
        out = _grids_u8(arr)
        assert out.dtype == np.uint8 and out.flags.c_contiguous
        assert np.array_equal(out, _old(arr))
    assert np.array_equal(_grids_u8(g.tolist()), _old(g))     # nested lists (env.py grids are lists of lists)


def test_grids_u8_rejects_single_map():
    try:
        _grids_u8(np.zeros((4, 4), int))
    except ValueError as e:
        assert "n_maps" in str(e)
    else:
        raise AssertionError("a 2-D grid is not a batch")


def test_shard_maps_without_group_is_everything():
    assert shard_maps(10) == (0, 10)
