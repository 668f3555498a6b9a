import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, ctypes as C
import torch.distributed as dist
from maaco_path_planing_b200 import MAACO, blocks_map, _lib
from maaco_path_planing_b200 import dist as dm
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
params = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
g = blocks_map(512, 0.2, seed=4000)
s = MAACO(g, 4096 * world, 100, rng_seed=4, device=local, group=dist.group.WORLD, verbose=False, max_cells=8192, **params)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
names = ['tours', 'exchange+best', 'pher+tau']
acc = {n: [] for n in names}
for it in range(1, 12):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    s._enqueue_iteration(it, events=ev)
    torch.cuda.synchronize()
    if it > 3:
        for i, n in enumerate(names): acc[n].append(ev[i].elapsed_time(ev[i + 1]))
if rank == 0:
    print('world', world, {n: round(float(np.mean(v)), 3) for n, v in acc.items()}, 'total', round(sum(float(np.mean(v)) for v in acc.values()), 3))
dist.destroy_process_group()
