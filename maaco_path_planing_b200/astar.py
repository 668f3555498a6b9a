"""AStarSolver -- drop-in for astar.AStarSolver (astar.py:10-101); `solve` runs the warp-cooperative
connector kernel (variant 0) and the statistics kernel.  `solve_batch` is the B200 addition: many
(start, target, avoid-set) queries in one launch."""
from __future__ import annotations

import numpy as np

from .gridmap import START_NODE_VAL, TARGET_NODE_VAL
from .helper import BasePathfinder


class AStarSolver(BasePathfinder):
    _variant = 0                                                            # connector semantics (0 = astar.py)
    def __init__(self, grid, turn_penalty_factor=0.1, safety_penalty_factor=0.05, min_safe_distance=1.5,
                 allow_diagonal_moves=True, restrict_diagonal_near_obstacle_policy=True,
                 diagonal_obstacle_penalty_value=1000.0, *, device=None, gridmap=None, engine=None):
        g = np.asarray(grid)
        s = np.argwhere(g == START_NODE_VAL)
        t = np.argwhere(g == TARGET_NODE_VAL)
        if not s.size > 0:
            raise ValueError("AStar: Start node not found in grid.")        # astar.py:19
        if not t.size > 0:
            raise ValueError("AStar: Target node not found in grid.")       # astar.py:20
        super().__init__(grid, tuple(s[0]), tuple(t[0]), turn_penalty_factor, safety_penalty_factor,
                         min_safe_distance, allow_diagonal_moves, restrict_diagonal_near_obstacle_policy,
                         diagonal_obstacle_penalty_value, device=device, gridmap=gridmap, engine=engine)
        self.astar_strictly_restricts_corners = self.restrict_diagonal_near_obstacle_policy

    def _avoid_bits(self, nodes_to_avoid):
        bits = np.zeros(self.engine.words, np.uint32)
        for r, c in nodes_to_avoid:
            j = int(r) * self.cols + int(c)
            bits[j >> 5] |= np.uint32(1 << (j & 31))
        return bits.view(np.int32)

    def solve_batch(self, starts, targets, avoid_sets=None):
        """[(path, popped_g)] for many queries in one kernel launch."""
        src = [self._cell(s) for s in starts]
        dst = [self._cell(t) for t in targets]
        bits = None
        if avoid_sets is not None:
            bits = np.stack([self._avoid_bits(a or ()) for a in avoid_sets])
        cells, ncell, g = self.engine.astar_batch(self._variant, src, dst, bits, self.allow_diagonal_moves,
                                                  self.astar_strictly_restricts_corners)
        cells, ncell, g = cells.cpu().numpy(), ncell.cpu().numpy(), g.cpu().numpy()
        return [(self._nodes(cells[i, :ncell[i]]), float(g[i])) for i in range(len(src))]

    def solve(self, start_node_override=None, target_node_override=None, nodes_to_avoid=None):   # astar.py:33
        s = start_node_override if start_node_override else self.start_node
        t = target_node_override if target_node_override else self.target_node
        inb = lambda n: 0 <= n[0] < self.rows and 0 <= n[1] < self.cols
        if not inb(s) or not inb(t):
            return self._calculate_stats_for_path([])                       # astar.py:37-39
        (path, g), = self.solve_batch([s], [t], [nodes_to_avoid] if nodes_to_avoid else None)
        if path and len(path) > 1:
            self.convergence_curve.append(g)                                # astar.py:70
        return self._calculate_stats_for_path(path)
