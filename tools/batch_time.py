"""Config-5 wave timing: M maps x 1024 ants x K passes in one MAACOBatch, for several ants-per-warp settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maaco_path_planing_b200 import blocks_map
from maaco_path_planing_b200.batch import MAACOBatch
P = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
grids = np.stack([blocks_map(256, 0.2, seed=5000 + i) for i in range(M)])
for apw in [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "8", "16", "32"])]:
    b = MAACOBatch(grids, 1024, K + 2, seeds=list(range(M)), ants_per_warp=apw, **P)
    b.run_iteration(1); b.run_iteration(2)
    torch.cuda.synchronize()
    s0 = b.total_steps()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(3, K + 3):
        b.run_iteration(it)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = b.total_steps() - s0
    print(f"maps {M} apw {apw}: {ms/K:.3f} ms/pass  {M*1024*K/ms/1e3:.1f} M evals/s  {st/ms/1e6:.2f} G ant-steps/s")
    b.close(); del b; torch.cuda.empty_cache()
