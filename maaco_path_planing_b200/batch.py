"""Batched sweeps over independent maps (BASELINE config 5): maps are independent units, so they are
sharded over the ranks with no data-path collective, and on each GPU several colonies run concurrently on
separate CUDA streams (a 1024-ant colony fills ~1/9 of a B200; the colony pass is enqueued without host
synchronisation, so interleaving the launches of a wave of solvers overlaps them)."""
from __future__ import annotations

from .dist import shard_range


def shard_maps(n_maps, group=None):
    """[lo, hi) of the maps this rank owns."""
    if group is None:
        return 0, n_maps
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_maps + world - 1) // world
    return min(n_maps, rank * per), min(n_maps, (rank + 1) * per)


def solve_maaco_batch(grids, num_ants, num_iterations, params, seeds=None, concurrent=12, group=None, device=None,
                      max_cells=None):
    """Run MAACO on every grid of `grids` (this rank's shard when `group` is given).

    Returns a list of (map_index, path, length, turns, convergence_curve) for the maps of this rank; results
    are identical to solving each map alone (same seeds => same Philox streams)."""
    import torch
    from .maaco import MAACO
    lo, hi = shard_maps(len(grids), group)
    out = []
    # the overlapped colonies fill the GPU together: pack the ants as one colony of that total size would be packed
    in_flight = max(1, min(concurrent, hi - lo)) * num_ants
    apw = 2 if in_flight <= 4096 else (4 if in_flight <= 8192 else 8)
    for w0 in range(lo, hi, concurrent):
        idx = list(range(w0, min(hi, w0 + concurrent)))
        solvers, streams = [], []
        for i in idx:
            g = grids[i]
            mc = max_cells or min(g.shape[0] * g.shape[1], 16 * (g.shape[0] + g.shape[1]))
            solvers.append(MAACO(g, num_ants, num_iterations, rng_seed=None if seeds is None else seeds[i],
                                 device=device, verbose=False, max_cells=mc, ants_per_warp=apw, **params))
            streams.append(torch.cuda.Stream(device=solvers[-1].device))
        cur = torch.cuda.current_stream(solvers[0].device)
        for st in streams:
            st.wait_stream(cur)
        for it in range(1, num_iterations + 1):                    # interleave the waves' launches
            for s, st in zip(solvers, streams):
                with torch.cuda.stream(st):
                    s._enqueue_iteration(it)
        for st in streams:
            st.synchronize()
        for i, s in zip(idx, solvers):
            s._iter_done = num_iterations
            path, length, turns = s.solve_path_planning()           # nothing left to enqueue: reads results back
            out.append((i, path, length, turns, list(s.convergence_curve_data)))
    return out


def solve_mpa_batch(grids, num_predators, num_iterations, params, seeds=None, group=None, device=None):
    """MPA over this rank's shard of the maps (one map at a time: an MPA iteration already fills the GPU with
    one warp per predator and needs a host round trip for its sorts)."""
    from .mpa import MPA
    lo, hi = shard_maps(len(grids), group)
    out = []
    for i in range(lo, hi):
        s = MPA(grids[i], num_predators, num_iterations, rng_seed=None if seeds is None else seeds[i], device=device,
                verbose=False, **params)
        res = s.solve_path_planning()
        out.append((i,) + tuple(res) + (list(s.convergence_curve_data),))
    return out
