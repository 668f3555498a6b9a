"""Config-3 fitness (4096 individuals x 5 waypoints, 512x512) for several numbers of search slots."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maaco_path_planing_b200 import _lib
if os.environ.get('MPP_SO'):
    _lib.SO_PATH = os.environ['MPP_SO']
from maaco_path_planing_b200 import GridMap, blocks_map
from maaco_path_planing_b200.engine import SearchEngine, make_policy
size, N, W = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 5
grid = blocks_map(size, 0.20, seed=3000 + size)
rng = np.random.default_rng(3)
free = np.flatnonzero(grid.ravel() != 1)
wps = torch.as_tensor(free[rng.integers(0, len(free), (N, W))].astype(np.int32), device="cuda")
pol = make_policy(0.3, 0.8, 1.8, 100.0)
gm = GridMap(grid)
for slots in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["3552", "1776", "888"])]:
    eng = SearchEngine(gm, n_slots=slots if slots > 0 else None)
    slots = eng.n_slots
    eng.waypoint_fitness(wps[:256], pol)
    torch.cuda.synchronize(); eng.counters.zero_()
    t0 = time.perf_counter()
    eng.waypoint_fitness(wps, pol)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e, r = eng.expansions()
    print(f"slots {slots}: {dt*1e3:.1f} ms  {e/dt/1e6:.1f} M exp/s  {dt/ (e/slots) *1e6:.2f} us/exp/slot  exp/eval {e/N:.0f}")
    del eng
    torch.cuda.empty_cache()
