import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
from maaco_path_planing_b200 import GridMap, blocks_map
from maaco_path_planing_b200.engine import SearchEngine, make_policy
for size, N in ((100, 1024), (256, 1024), (512, 1024), (512, 4096)):
    grid = blocks_map(size, 0.2, seed=3000 + size)
    rng = np.random.default_rng(1)
    free = np.flatnonzero(grid.ravel() != 1)
    wps = free[rng.integers(0, len(free), (N, 5))].astype(np.int32)
    eng = SearchEngine(GridMap(grid))
    pol = make_policy(0.3, 0.8, 1.8, 100.0)
    eng.waypoint_fitness(wps[:64], pol); torch.cuda.synchronize()
    eng.counters.zero_()
    t0 = time.time(); cells, ncell, stats = eng.waypoint_fitness(wps, pol); torch.cuda.synchronize(); dt = time.time() - t0
    e, r = eng.expansions()
    nc = ncell.cpu().numpy()
    print(f'size={size} N={N} time={dt*1e3:.1f} ms evals/s={N/dt:.0f} valid={(nc>0).sum()} meanlen={nc[nc>0].mean():.0f} expansions={e} ({e/N:.0f}/eval) exp/s={e/dt/1e6:.1f}M relax={r} heap_cap={eng.heap_cap} slots={eng._scratch_key}', flush=True)
    print('   queue: ring pushes, heap pushes =', eng.queue_stats())
