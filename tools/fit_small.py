import sys
sys.path.insert(0, '.')
import numpy as np, torch
from maaco_path_planing_b200 import GridMap, blocks_map
from maaco_path_planing_b200.engine import SearchEngine, make_policy
size, N = 256, 1024
grid = blocks_map(size, 0.2, seed=3000 + size)
rng = np.random.default_rng(1)
free = np.flatnonzero(grid.ravel() != 1)
eng = SearchEngine(GridMap(grid))
pol = make_policy(0.3, 0.8, 1.8, 100.0)
for k in range(2):
    wps = free[rng.integers(0, len(free), (N, 5))].astype(np.int32)
    eng.waypoint_fitness(wps, pol)
torch.cuda.synchronize()
print('ok', eng.expansions())
