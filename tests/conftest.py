import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CUDA library and the oracle are built (cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "params" in d:
        d["params"] = json.loads(str(d["params"]))
    return d


MAACO_CASES = ["fig7", "fig13", "blocks64", "near_roulette", "near_alpha2", "rect24x40", "aligned_row", "reverse_diag"]
# main.py's own parameters (50 ants x 100 iterations) on every 20x20 demo map + grid_map_from_image_data5 (256x256)
MAACO_DEFAULT_CASES = ["fig7_default", "fig13_default", "image1_default", "image2_default", "image3_default", "image5"]
MAACO_DEFAULT = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2,
                     k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
