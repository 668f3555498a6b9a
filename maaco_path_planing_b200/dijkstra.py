"""DijkstraSolver -- drop-in for dijkstra.DijkstraSolver (dijkstra.py:10-97): the connector kernel with a
zero heuristic (variant 2), i.e. extract-min over (g, r, c)."""
from __future__ import annotations

import numpy as np

from .astar import AStarSolver
from .gridmap import START_NODE_VAL, TARGET_NODE_VAL
from .helper import BasePathfinder


class DijkstraSolver(AStarSolver):
    _variant = 2

    def __init__(self, grid, turn_penalty_factor=0.1, safety_penalty_factor=0.05, min_safe_distance=1.5,
                 allow_diagonal_moves=True, restrict_diagonal_near_obstacle_policy=True,
                 diagonal_obstacle_penalty_value=1000.0, *, device=None, gridmap=None, engine=None):
        g = np.asarray(grid)
        s = np.argwhere(g == START_NODE_VAL)
        t = np.argwhere(g == TARGET_NODE_VAL)
        if not s.size > 0:
            raise ValueError("Dijkstra: Start node not found.")                # dijkstra.py:19
        if not t.size > 0:
            raise ValueError("Dijkstra: Target node not found.")               # dijkstra.py:20
        BasePathfinder.__init__(self, grid, tuple(s[0]), tuple(t[0]), turn_penalty_factor, safety_penalty_factor,
                                min_safe_distance, allow_diagonal_moves, restrict_diagonal_near_obstacle_policy,
                                diagonal_obstacle_penalty_value, device=device, gridmap=gridmap, engine=engine)
        self.astar_strictly_restricts_corners = self.restrict_diagonal_near_obstacle_policy
        self.dijkstra_strictly_restricts_corners = self.restrict_diagonal_near_obstacle_policy
