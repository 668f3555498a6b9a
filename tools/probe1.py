import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import pyoracle as O
from maaco_path_planing_b200 import MAACO, blocks_map
params = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
g = blocks_map(512, 0.2, seed=4000)
N, K, seed = 4096, 100, 4
for lpa in (32, 8):
    dev = MAACO(g, N, K, rng_seed=seed, device=0, verbose=False, lanes_per_ant=lpa, **params)
    orc = O.MaacoOracle(g, N, K, seed=seed, threads=0, **params) if lpa == 32 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(1, 4):
        t0 = time.time()
        ev0.record(); dev.run_iteration(it); ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        nc, ln, tn, cells = dev.last_tours()
        msg = ''
        if orc is not None:
            t1 = time.time(); ocells, onc, oln, otn, _ = orc.iterate(it); tor = time.time() - t1
            ok = np.array_equal(nc, onc) and np.array_equal(ln, oln) and np.array_equal(tn, otn) and all(np.array_equal(cells[a,:nc[a]], ocells[a,:onc[a]]) for a in range(N))
            okt = np.array_equal(dev.pheromone_matrix.ravel(), orc.tau)
            msg = f'parity tours={ok} tau={okt} oracle_s={tor:.2f}'
        print(f'lpa={lpa} it={it} ms={ms:.3f} succ={(nc>0).sum()} meanlen={nc[nc>0].mean():.0f} maxlen={nc.max()} steps={dev.total_steps()} {msg}', flush=True)
    # timing breakdown
    import ctypes as C
    from maaco_path_planing_b200 import _lib
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    L = _lib.lib(); stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for it in range(4, 8):
        q0 = dev._calculate_adaptive_q0(it)
        ev[0].record()
        L.mpp_maaco_tours(dev.map.handle, _lib.ptr(dev._tau), _lib.ptr(dev._E0), _lib.ptr(dev._E1), it, q0, dev.alpha, N, 0, C.c_uint64(seed), _lib.ptr(dev._visitT), _lib.ptr(dev._cells), dev.max_cells, _lib.ptr(dev._n_cells), _lib.ptr(dev._length), _lib.ptr(dev._turns), _lib.ptr(dev._steps), lpa, stream)
        ev[1].record()
        L.mpp_maaco_best(_lib.ptr(dev._length), _lib.ptr(dev._turns), _lib.ptr(dev._n_cells), _lib.ptr(dev._cells), dev.max_cells, N, dev.Q, it, _lib.ptr(dev._state), _lib.ptr(dev._best_cells), _lib.ptr(dev._deposit), _lib.ptr(dev._log), stream)
        ev[2].record()
        L.mpp_maaco_pheromone(dev.map.handle, _lib.ptr(dev._tau), _lib.ptr(dev._visitT), _lib.ptr(dev._deposit), N, dev.rho, _lib.ptr(dev._state), 1, stream)
        ev[3].record(); torch.cuda.synchronize()
        print(f'  it={it} tours={ev[0].elapsed_time(ev[1]):.3f} best={ev[1].elapsed_time(ev[2]):.3f} pher={ev[2].elapsed_time(ev[3]):.3f} ms')
    del dev
