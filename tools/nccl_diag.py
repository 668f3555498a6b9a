"""torchrun helper: the NCCL (all-gather) exchange of the sharded colony against the oracle, pass by pass, without and
with a forced overflow of the exchange buffer; prints the first deviating pass on rank 0."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ["MPP_P2P"] = "0"
import pyoracle as O
from maaco_path_planing_b200 import MAACO, blocks_map
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
P = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
g = blocks_map(96, 0.2, seed=11)
N, K, seed = 64 * world, 6, 9
def run(tag, force):
    dev = MAACO(g, N, K, rng_seed=seed, device=local, group=dist.group.WORLD, verbose=False, **P)
    if force:
        dev._cap = dev._cap_max = 0
    orc = O.MaacoOracle(g, N, K, seed=seed, **P)
    bad = None
    for it in range(1, K + 1):
        dev.run_iteration(it)
        _, onc, oln, otn, _ = orc.iterate(it)
        if force and it < K:
            continue                                  # (settling every pass would hide the lagged confirmation)
        nc, ln, tn = dev.last_results()
        tau = dev.pheromone_matrix.ravel()
        okr = np.array_equal(nc, onc) and np.array_equal(ln, oln) and np.array_equal(tn, otn)
        okt = np.array_equal(tau, orc.tau)
        if not (okr and okt) and bad is None:
            d = np.flatnonzero(tau != orc.tau)
            bad = (it, okr, okt, len(d), d[:6].tolist(), [float(tau[i] - orc.tau[i]) for i in d[:3]])
    flag = torch.tensor([0 if bad is None else 1], device="cuda"); dist.all_reduce(flag)
    if rank == 0 or bad is not None:
        print(f"[{tag}] rank {rank}/{world}: {'ok' if bad is None else 'first bad pass, results ok, tau ok, #cells, cells, diffs = ' + str(bad)}; rewinds {getattr(dev, 'exchange_rewinds', 0)}; caps {dev._cap} {dev._cap_max}", flush=True)
    dist.barrier()
run("nccl, normal", False)
run("nccl, forced overflow", True)
dist.destroy_process_group()
