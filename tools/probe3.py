import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch, ctypes as C
from maaco_path_planing_b200 import MAACO, blocks_map, _lib
params = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
g = blocks_map(512, 0.2, seed=4000)
for N in (4096, 32768):
  for lpa in (32, 16, 8):
    dev = MAACO(g, N, 100, rng_seed=4, device=0, verbose=False, lanes_per_ant=lpa, max_cells=4096, **params)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tt=[];pp=[]
    for it in range(1, 8):
        ev[0].record(); dev._enqueue_tours(it, stream); ev[1].record(); dev._enqueue_best(it, stream); ev[2].record(); dev._enqueue_pheromone(stream); ev[3].record(); torch.cuda.synchronize()
        if it>2: tt.append(ev[0].elapsed_time(ev[1])); pp.append(ev[2].elapsed_time(ev[3]))
    print(f'N={N} lpa={lpa} tours={np.mean(tt):.3f} ms pher={np.mean(pp):.3f} ms steps/pass={dev.total_steps()/7:.0f} -> {dev.total_steps()/7/np.mean(tt)/1e6:.2f} Gsteps/s', flush=True)
    del dev; torch.cuda.empty_cache()
