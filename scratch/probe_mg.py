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
names = ['tours', 'ag_results', 'a2a_visit', 'memset', 'best', 'pher', 'tau_copy', 'ag_tau']
acc = {n: [] for n in names}
for it in range(1, 12):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    nl, off = s.n_local, s.ant_offset
    ev[0].record(); s._enqueue_tours(it, stream)
    ev[1].record(); dm.exchange_results(s._result, s._result[off:off + nl].clone(), s.group)
    ev[2].record(); dm.exchange_visit_slices(s._visit_recv, s._visit_local, s.group)
    ev[3].record(); s._visit_local.zero_()
    ev[4].record(); s._enqueue_best(it, stream)
    ev[5].record(); s._enqueue_pheromone(stream)
    ev[6].record(); wn32 = s.words_per_rank * 32; s._tau_slice.copy_(s._tau[s.rank * wn32:(s.rank + 1) * wn32])
    ev[7].record(); dm.gather_tau(s._tau, s._tau_slice, s.group)
    ev[8].record(); torch.cuda.synchronize()
    if it > 3:
        for i, n in enumerate(names): acc[n].append(ev[i].elapsed_time(ev[i + 1]))
if rank == 0:
    print('world', world, {n: round(float(np.mean(v)), 3) for n, v in acc.items()}, 'total', round(sum(float(np.mean(v)) for v in acc.values()), 3))
dist.destroy_process_group()
