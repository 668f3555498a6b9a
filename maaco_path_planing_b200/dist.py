"""Sharding helpers for the one place the hot path has a real exchange step: the MAACO colony.

Ants are independent given tau (MAACO.py:340-342), so rank g constructs ants
[g*N/G, (g+1)*N/G); the pheromone update (MAACO.py:304-311) needs every ant's visited cells in
*global ant order* to stay bit-exact.  Per pass:

  1. ONE all-gather of a per-rank buffer holding the 16-byte per-ant results (length, n_cells, turns) and the tours
     as 1-byte move codes (~0.5 KB/ant; sized from the totals seen two passes earlier, csrc/mpp_maaco.cu
     "Sharded-colony exchange") -> replicated best scan; every rank replays all ants into the visited-set slabs of
     ITS slice of tile rows and updates the pheromone of that slice with all N ants in global order, so tau is
     bit-identical to the single-GPU run;
  2. all-gather of the updated tau slices (replicated tau for the next colony pass).

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


def padded_tile_rows(tile_rows, world):
    """Tile rows (32 cell rows each) of the map, padded so that they split evenly over the ranks."""
    return ((tile_rows + world - 1) // world) * world


def exchange_buffers(buf_all, buf_local, group):
    """buf_local: this rank's exchange buffer [seg] uint8 -> buf_all [G*seg] in rank order."""
    dist.all_gather_into_tensor(buf_all, buf_local, group=group)


def gather_tau(tau_full, tau_slice, group):
    """tau_slice: this rank's slice of whole tile rows -> tau_full (padded, replicated)."""
    dist.all_gather_into_tensor(tau_full, tau_slice, group=group)
