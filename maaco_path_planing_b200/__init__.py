"""maaco_path_planing_b200 -- B200-native population evaluation for the grid path planners of
dvnam1605/MAACO-path-planing (MAACO tours + pheromone update, A*-connector fitness for PSO/GA,
MPA path reconstruction).  Python host code over hand-written sm_100a kernels behind a C ABI
(include/mpp.h); no CPU fallback.  Drop-in module names live in ``maaco_path_planing_b200/dropin``.
"""
from ._lib import MppError, SO_PATH, lib  # noqa: F401
from .gridmap import (FREE_SPACE, OBSTACLE, START_NODE_VAL, TARGET_NODE_VAL, GridMap, blocks_map)  # noqa: F401
from .maaco import MAACO  # noqa: F401

__all__ = ["MAACO", "GridMap", "blocks_map", "MppError", "FREE_SPACE", "OBSTACLE", "START_NODE_VAL",
           "TARGET_NODE_VAL"]
