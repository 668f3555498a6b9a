"""Times the ranking + tour kernels of BASELINE config 4 (4096 ants, 512x512) for a given libmpp build.
    MPP_SO=/path/to/lib.so python tools/tour_time.py [ants] [size] [passes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maaco_path_planing_b200 import _lib
if os.environ.get("MPP_SO"):
    _lib.SO_PATH = os.environ["MPP_SO"]
from maaco_path_planing_b200 import MAACO, blocks_map
P = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
ants = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
g = blocks_map(size, 0.2, seed=4000)
s = MAACO(g, ants, K + 8, rng_seed=4, verbose=False, **P)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(1, 4):
    s.run_iteration(it)
torch.cuda.synchronize()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
for k in range(K):
    flush.fill_(k)
    s._enqueue_iteration(4 + k, events=ev[k])
torch.cuda.synchronize()
f = lambda a, b: sum(e[a].elapsed_time(e[b]) for e in ev) / K
print(f"{os.environ.get('MPP_SO','default')}: pass {f(0,3):.4f} ms  rank {f(0,4):.4f}  tours {f(4,1):.4f}  best {f(1,2):.4f}  pheromone {f(2,3):.4f}  steps/pass {s.total_steps()/(K+3):.0f}")
