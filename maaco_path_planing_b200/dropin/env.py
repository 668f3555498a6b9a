"""Drop-in for the reference module `env` (env.py): the four cell constants and the grid-map
interface (list-of-lists / ndarray of 0/1/2/3).  The reference's literal demo maps are data, not
code, and are not duplicated here: `load_reference_grids(path_to_reference_env_py)` reads them from
a reference checkout, `blocks(n, frac, seed)` generates the synthetic benchmark maps."""
import _bootstrap  # noqa: F401
import numpy as np  # noqa: E402

from maaco_path_planing_b200.gridmap import (FREE_SPACE, OBSTACLE, START_NODE_VAL, TARGET_NODE_VAL,  # noqa: F401,E402
                                             blocks_map as blocks)


def load_reference_grids(env_py_path):
    """Evaluate the grid literals of a reference env.py and return {name: list-of-lists}."""
    ns = {"np": np}
    with open(env_py_path) as f:
        exec(compile(f.read(), env_py_path, "exec"), ns)
    return {k: v for k, v in ns.items() if k.startswith("grid_")}
