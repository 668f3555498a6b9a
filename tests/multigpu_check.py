"""torchrun helper: a colony sharded over WORLD_SIZE GPUs must reproduce the single-colony oracle
trajectory bit-for-bit (same global ant ids => same Philox streams, ordered deposit).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import pyoracle as O
    from maaco_path_planing_b200 import MAACO, blocks_map
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    params = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9,
                  q0_initial=0.5, C0_initial_pheromone=0.1)
    g = blocks_map(96, 0.2, seed=11)
    N, K, seed = 64 * world, 6, 9
    dev = MAACO(g, N, K, rng_seed=seed, device=local, group=dist.group.WORLD, verbose=False, **params)
    orc = O.MaacoOracle(g, N, K, seed=seed, **params)
    for it in range(1, K + 1):
        dev.run_iteration(it)
        nc, ln, tn = dev.last_results()
        _, onc, oln, otn, _ = orc.iterate(it)
        assert np.array_equal(nc, onc) and np.array_equal(ln, oln) and np.array_equal(tn, otn), f"rank {rank} it {it}"
        assert np.array_equal(dev.pheromone_matrix.ravel(), orc.tau), f"rank {rank} tau it {it}"
    dev2 = MAACO(g, N, K, rng_seed=seed, device=local, group=dist.group.WORLD, verbose=False, **params)
    path, length, turns = dev2.solve_path_planning()
    orc2 = O.MaacoOracle(g, N, K, seed=seed, **params)
    opath, olen, oturns = orc2.solve()
    assert [r * 96 + c for r, c in path] == list(opath) and length == olen and turns == oturns, f"rank {rank} solve"
    # an exchange buffer that is far too small: the device latch stops the colony, the host makes room and repeats
    # (the all-gather exchange: with peer memory nothing is sized, so nothing can overflow)
    p2p_env = os.environ.get("MPP_P2P")
    os.environ["MPP_P2P"] = "0"
    dev3 = MAACO(g, N, K, rng_seed=seed, device=local, group=dist.group.WORLD, verbose=False, **params)
    if p2p_env is None:
        del os.environ["MPP_P2P"]
    else:
        os.environ["MPP_P2P"] = p2p_env
    dev3._cap = dev3._cap_max = 0
    orc3 = O.MaacoOracle(g, N, K, seed=seed, **params)
    for it in range(1, K + 1):
        dev3.run_iteration(it)
        orc3.iterate(it)
    assert np.array_equal(dev3.pheromone_matrix.ravel(), orc3.tau), f"rank {rank} tau after rewind"
    assert getattr(dev3, "exchange_rewinds", 0) >= 1, "the tiny exchange buffer should have overflowed"
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok: world={world}, {N} ants, {K} passes bit-exact vs oracle (peer memory: {dev._p2p is not None})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
