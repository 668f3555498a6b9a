"""Drop-in for the reference module `env` (env.py): the four cell constants (env.py:4-7) and the six demo
maps `main.py:9-17` imports (env.py:10-371), as list-of-lists of 0/1/2/3 exactly like the reference defines
them (start/target are marked by the caller, main.py:27-32, except where the map itself carries 2/3).
The maps are data, not code: they are read from `maaco_path_planing_b200/data/env_grids.npz`
(written by tests/golden/make_golden.py from the reference's literals).  `blocks(n, frac, seed)` generates the
synthetic benchmark maps; `load_reference_grids(path)` reads the literals of any other env.py."""
import os as _os

import _bootstrap  # noqa: F401
import numpy as np  # noqa: E402

from maaco_path_planing_b200.gridmap import (FREE_SPACE, OBSTACLE, START_NODE_VAL, TARGET_NODE_VAL,  # noqa: F401,E402
                                             blocks_map as blocks)

_DATA = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "data", "env_grids.npz")
GRID_NAMES = ("grid_fig7_layout_data", "grid_map_fig13_base_data", "grid_map_from_image_data",
              "grid_map_from_image_data2", "grid_map_from_image_data3", "grid_map_from_image_data5")


def _load():
    with np.load(_DATA, allow_pickle=False) as z:
        return {k: z[k].astype(int).tolist() for k in GRID_NAMES}     # list-of-lists of Python ints, like env.py


globals().update(_load())


def load_reference_grids(env_py_path):
    """Evaluate the grid literals of a reference env.py and return {name: list-of-lists}."""
    ns = {"np": np}
    with open(env_py_path) as f:
        exec(compile(f.read(), env_py_path, "exec"), ns)
    return {k: v for k, v in ns.items() if k.startswith("grid_")}
