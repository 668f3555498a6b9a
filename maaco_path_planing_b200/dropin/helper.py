"""Drop-in for the reference module `helper` (helper.py:8-161)."""
import _bootstrap  # noqa: F401
from maaco_path_planing_b200.helper import (BasePathfinder, count_turns, distance_euclidean,  # noqa: F401,E402
                                            get_valid_neighbors, heuristic_euclidean, is_valid_and_not_obstacle)
