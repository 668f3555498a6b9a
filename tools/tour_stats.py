"""Per-ant tier mix / cycles of the tour kernel (needs a build with -DMPP_TOUR_STATS; MPP_SO=/path/to/lib.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maaco_path_planing_b200 import _lib
_lib.SO_PATH = os.environ["MPP_SO"]
from maaco_path_planing_b200 import MAACO, blocks_map
P = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
ants = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = blocks_map(512, 0.2, seed=4000)
s = MAACO(g, ants, 16, rng_seed=4, verbose=False, **P)
for it in range(1, 6):
    s.run_iteration(it)
torch.cuda.synchronize()
mv = s._moves.view(ants, s.max_cells)[:, -32:].cpu().numpy().copy().view(np.int32)[:, :7]
t1, t2, t3, sl, cyc, steps, ok = [mv[:, i].astype(np.int64) for i in range(7)]
print("ants", ants, "steps total", steps.sum(), "tiers %", 100 * t1.sum() / steps.sum(), 100 * t2.sum() / steps.sum(), 100 * t3.sum() / steps.sum(), "slides/step", sl.sum() / steps.sum())
order = np.argsort(-cyc)[:12]
print("slowest ants: steps t1 t2 t3 slides cycles cyc/step ok")
for a in order:
    print(a, steps[a], t1[a], t2[a], t3[a], sl[a], cyc[a], round(cyc[a] / max(1, steps[a])), ok[a])
print("median cyc/step", np.median(cyc[steps > 50] / steps[steps > 50]), "mean steps", steps.mean(), "max", steps.max())
# crude per-tier cost: least squares cycles ~ a*t1 + b*t2 + c*t3 + d*slides
Aa = np.stack([t1, t2, t3, sl], 1).astype(float)
coef, *_ = np.linalg.lstsq(Aa, cyc.astype(float), rcond=None)
print("fit cycles per: tier1 %.0f tier2 %.0f tier3 %.0f slide %.0f" % tuple(coef))
