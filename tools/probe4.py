import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import pyoracle as O
from maaco_path_planing_b200 import GridMap, blocks_map
from maaco_path_planing_b200.engine import SearchEngine, make_policy
size = 512
grid = blocks_map(size, 0.2, seed=3000 + size)
rng = np.random.default_rng(1)
free = np.flatnonzero(grid.ravel() != 1)
N = 256
wps = free[rng.integers(0, len(free), (N, 5))].astype(np.int32)
eng = SearchEngine(GridMap(grid))
pol = make_policy(0.3, 0.8, 1.8, 100.0)
eng.waypoint_fitness(wps[:8], pol); torch.cuda.synchronize()
# per-individual expansions from the oracle
ex = []
for i in range(N):
    _, _, _, e = O.waypoint_fitness(grid, wps[i:i+1], 0.3, 0.8, 1.8, 100.0)
    ex.append(e)
ex = np.array(ex)
print('expansions per individual: mean %.0f max %d p90 %.0f' % (ex.mean(), ex.max(), np.percentile(ex, 90)))
# time single individuals: the max one and a typical one
for i in (int(ex.argmax()), int(np.argsort(ex)[N // 2])):
    t0 = time.time(); eng.waypoint_fitness(wps[i:i+1], pol); torch.cuda.synchronize(); dt = time.time() - t0
    print(f'individual {i}: expansions {ex[i]} time {dt*1e3:.1f} ms -> {dt/ex[i]*1e9:.0f} ns/expansion (1 warp alone)')
for n in (32, 256):
    t0 = time.time(); eng.waypoint_fitness(wps[:n], pol); torch.cuda.synchronize(); dt = time.time() - t0
    print(f'N={n}: {dt*1e3:.1f} ms; max exp in batch {ex[:n].max()} -> {dt/ex[:n].max()*1e9:.0f} ns per expansion of the longest chain')
