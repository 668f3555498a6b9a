"""helper -- the reference's shared path primitives (helper.py:8-161) over the B200 engine.

`BasePathfinder` keeps the reference attributes (grid, rows, cols, start_node, target_node,
obstacle_nodes, the six policy scalars, convergence_curve) and computes `_calculate_stats_for_path`
with the CUDA statistics kernel.  The small pure functions (distance, neighbours, turn count) are
host utilities with the reference's semantics; the hot versions live in the kernels.
"""
from __future__ import annotations

import math

import numpy as np

from .engine import SearchEngine, make_policy
from .gridmap import OBSTACLE, GridMap

INF = float("inf")


def distance_euclidean(node1, node2):                                   # helper.py:8-12
    dr = node1[0] - node2[0]
    dc = node1[1] - node2[1]
    return math.sqrt(dr ** 2 + dc ** 2)


heuristic_euclidean = distance_euclidean                               # helper.py:55-56


def is_valid_and_not_obstacle(r, c, grid, rows, cols):                  # helper.py:14-16
    return 0 <= r < rows and 0 <= c < cols and grid[r, c] != OBSTACLE


def get_valid_neighbors(r, c, grid, rows, cols, allow_diagonal_moves=True, restrict_diagonal_corner_cutting=True,
                        exclude_nodes=None):                            # helper.py:18-53
    exclude_nodes = exclude_nodes or set()
    moves = [(0, 1, False), (0, -1, False), (1, 0, False), (-1, 0, False)]
    if allow_diagonal_moves:
        moves += [(1, 1, True), (1, -1, True), (-1, 1, True), (-1, -1, True)]
    out = []
    for dr, dc, diag in moves:
        nr, nc = r + dr, c + dc
        if not is_valid_and_not_obstacle(nr, nc, grid, rows, cols) or (nr, nc) in exclude_nodes:
            continue
        if diag and restrict_diagonal_corner_cutting and (
                not is_valid_and_not_obstacle(r + dr, c, grid, rows, cols)
                or not is_valid_and_not_obstacle(r, c + dc, grid, rows, cols)):
            continue
        out.append((nr, nc))
    return out


def count_turns(path):                                                  # helper.py:58-65
    if len(path) < 3:
        return 0
    turns = 0
    for i in range(len(path) - 2):
        d1 = (path[i + 1][0] - path[i][0], path[i + 1][1] - path[i][1])
        d2 = (path[i + 2][0] - path[i + 1][0], path[i + 2][1] - path[i + 1][1])
        turns += d1 != d2
    return turns


class BasePathfinder:                                                   # helper.py:115-161
    def __init__(self, grid, start_node, target_node, turn_penalty_factor, safety_penalty_factor, min_safe_distance,
                 allow_diagonal_moves, restrict_diagonal_near_obstacle_policy, diagonal_obstacle_penalty_value,
                 *, device=None, gridmap=None, engine=None):
        self.grid = np.array(grid, dtype=int)
        self.rows, self.cols = self.grid.shape
        self.start_node = (int(start_node[0]), int(start_node[1]))
        self.target_node = (int(target_node[0]), int(target_node[1]))
        self.obstacle_nodes = np.argwhere(self.grid == OBSTACLE)
        self.turn_penalty_factor = turn_penalty_factor
        self.safety_penalty_factor = safety_penalty_factor
        self.min_safe_distance = min_safe_distance
        self.allow_diagonal_moves = allow_diagonal_moves
        self.restrict_diagonal_near_obstacle_policy = restrict_diagonal_near_obstacle_policy
        self.diagonal_obstacle_penalty_value = diagonal_obstacle_penalty_value
        self.convergence_curve = []
        self.map = gridmap if gridmap is not None else GridMap(self.grid, device=device)
        self.engine = engine if engine is not None else SearchEngine(self.map)
        self.policy = make_policy(turn_penalty_factor, safety_penalty_factor, min_safe_distance,
                                  diagonal_obstacle_penalty_value, restrict_diagonal_near_obstacle_policy,
                                  allow_diagonal_moves, mode=0)

    def _cell(self, node):
        return int(node[0]) * self.cols + int(node[1])

    def _nodes(self, cells):
        return [(int(c) // self.cols, int(c) % self.cols) for c in cells]

    def _calculate_stats_for_path(self, path):                          # helper.py:138-147
        return self.engine.stats_of_path(list(path), self.policy)

    def plot_convergence_curve(self, title_prefix="Algorithm"):         # helper.py:149-161 (plotting: out of scope)
        try:
            import matplotlib.pyplot as plt
        except Exception:
            print(f"matplotlib not available; {title_prefix} convergence data is in .convergence_curve")
            return
        data = [f for f in self.convergence_curve if f is not None and f != INF]
        if data:
            plt.figure(); plt.plot(data); plt.title(f"{title_prefix} Convergence Curve")
            plt.xlabel("Iteration / Evaluation"); plt.ylabel("Best Fitness"); plt.grid(True)
        else:
            print(f"No valid convergence data to plot for {title_prefix}.")
