"""torchrun helper: per-phase CUDA-event times of the sharded colony pass (weak: 4096 ants per GPU; --strong: 4096 in all).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/shard_time.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from maaco_path_planing_b200 import MAACO, blocks_map

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
P = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
n = 4096 if "--strong" in sys.argv else 4096 * world
g = blocks_map(512, 0.2, seed=4000)
s = MAACO(g, n, 64, rng_seed=4, device=local, group=dist.group.WORLD, verbose=False, **P)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(1, 5):
    s.run_iteration(it)
acc = {}
K = 8
for it in range(5, 5 + K):
    flush.fill_(it)
    dist.barrier()
    s._phase_log = []
    s.run_iteration(it)
    s._settle()
    log = s._phase_log
    for (n0, e0), (n1, e1) in zip(log[:-1], log[1:]):
        acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / K
    acc["pass"] = acc.get("pass", 0.0) + log[0][1].elapsed_time(log[-1][1]) / K
s._phase_log = None
for r in range(world):
    dist.barrier()
    if r == rank and (rank == 0 or rank == world - 1):
        print(f"rank {rank}/{world} ants {n} p2p={s._p2p is not None}: " + "  ".join(f"{k} {v*1e3:.0f}us" for k, v in acc.items()), flush=True)
dist.destroy_process_group()
