"""Sharding helpers for the one place the hot path has a real exchange step: the MAACO colony.

Ants are independent given tau (MAACO.py:340-342), so rank g constructs ants
[g*N/G, (g+1)*N/G); the pheromone update (MAACO.py:304-311) needs every ant's visited cells in
*global ant order* to stay bit-exact.  The exchange is therefore:

  1. all-gather of the 16-byte per-ant results (length, n_cells, turns) -> replicated best scan;
  2. the visited sets, one of two ways:
     "moves" (default): all-gather of the tours as 1-byte move codes (~0.8 KB/ant); every rank replays
        all ants and sets the visited bits that fall into ITS slice of bitmap words;
     "dense": all-to-all of visited-bitmap word slices -- rank g receives, from every source rank s, the
        words [g*Wn, (g+1)*Wn) of s's ants ([Wn][N_local], contiguous in the word-major layout);
     either way rank g updates only the cells of its slice, with all N ants in global order
     (segments = source ranks), so tau is bit-identical to the single-GPU run;
  3. all-gather of the updated tau slices (replicated tau for the next colony pass).

Everything here is plain torch / torch.distributed on whatever device the tensors live on, so the
same code runs under NCCL on B200s and under gloo on CPU (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total, world, rank):
    """Contiguous equal shards [lo, hi) of n_total units; n_total must divide evenly."""
    if n_total % world:
        raise ValueError(f"{n_total} units do not shard evenly over {world} ranks")
    per = n_total // world
    return rank * per, (rank + 1) * per


def padded_words(n_cells, world):
    """Bitmap words per ant, padded so that the word axis splits evenly over the ranks."""
    w = (n_cells + 31) // 32
    return ((w + world - 1) // world) * world


def _all_to_all(out, inp, group):
    if dist.get_backend(group) == "gloo":
        # gloo has no all_to_all_single: emulate with an all-gather of every rank's full buffer
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        bufs = [torch.empty_like(inp) for _ in range(world)]
        dist.all_gather(bufs, inp, group=group)
        n = inp.numel() // world
        for s in range(world):
            out[s * n:(s + 1) * n].copy_(bufs[s][rank * n:(rank + 1) * n])
    else:
        dist.all_to_all_single(out, inp, group=group)


def exchange_results(result_all, result_local, group):
    """result_local: [N_local, 2] int64 view of mpp_ant_result -> result_all [N, 2] in rank order."""
    dist.all_gather_into_tensor(result_all, result_local, group=group)


def exchange_visit_slices(visit_recv, visit_local, group):
    """visit_local: flat [W_pad * N_local] (word-major) -> visit_recv flat [G][Wn][N_local]."""
    _all_to_all(visit_recv, visit_local, group)


def gather_tau(tau_full, tau_slice, group):
    """tau_slice: this rank's [Wn*32] cells -> tau_full [W_pad*32] (replicated)."""
    dist.all_gather_into_tensor(tau_full, tau_slice, group=group)


def exchange_moves(packed_all, packed_local, group):
    """packed_local: this rank's move-code buffer [cap] uint8 -> packed_all [G*cap] in rank order."""
    dist.all_gather_into_tensor(packed_all, packed_local, group=group)
