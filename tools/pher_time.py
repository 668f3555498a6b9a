"""Times mpp_maaco_pheromone alone on the slabs of a real config-4 pass: as is, and with no depositing ant (fixed cost)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maaco_path_planing_b200 import _lib
if os.environ.get("MPP_SO"):
    _lib.SO_PATH = os.environ["MPP_SO"]
from maaco_path_planing_b200 import MAACO, blocks_map
P = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2, k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)
g = blocks_map(512, 0.2, seed=4000)
s = MAACO(g, 4096, 16, rng_seed=4, verbose=False, **P)
for it in (1, 2, 3):
    s.run_iteration(it)
torch.cuda.synchronize()
L = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
NOFLUSH = False
def run(tag, it=3):
    ms = []
    for k in range(6):
        # re-create the touched bitmap of parity it&1 (the update cleared the other parity only) -- it is still intact
        if not NOFLUSH:
            flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.mpp_maaco_pheromone(s._maps, C.byref(s._colony), _lib.ptr(s._slabs), _lib.ptr(s._touched), 4096, 0, s.tile_rows, 0.1, it, 0, None, 0, st), "pher")
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(tag, " ".join(f"{x*1e3:.1f}" for x in ms), "us")
run("as is (L2 flushed before)")
ok = s._okbits.clone()
s._okbits.zero_()
run("no depositing ant")
s._okbits.copy_(ok)
# only the start tile's ants
def dump_prof(tag, top=6):
    try:
        fn = L.mpp_debug_pher_prof
    except AttributeError:
        return
    import numpy as np
    out = np.zeros(8192 * 8, dtype=np.uint64)
    fn(out.ctypes.data_as(C.c_void_p))
    o = out.reshape(8192, 8).astype(np.int64)
    t0 = o[:, 0].min()
    start, end, hits, rounds = o[:, 0] - t0, o[:, 1] - t0, o[:, 2], o[:, 3]
    dur = end - start
    print(f"  [{tag}] span {end.max()/1e3:.1f} us; warp starts median {np.median(start)/1e3:.1f} max {start.max()/1e3:.1f}; sum warp time {dur.sum()/1e6:.1f} ms")
    for i in np.argsort(-end)[:top]:
        cta, w = divmod(i, 8)
        print(f"    cta {cta} (b {cta >> 2} rg {cta & 3}) w {w}: start {start[i]/1e3:.1f} end {end[i]/1e3:.1f} dur {dur[i]/1e3:.1f} us, rounds {rounds[i]} hits {hits[i]}; extra {o[i,4]} {o[i,5]} {o[i,6]} {o[i,7]}")
dump_prof("real pass data")

# ---- synthetic slabs: one tile touched by every ant; what does a round cost as a function of the bit pattern? ----
import numpy as np
TCn = 16
def synth(tag, words_fn, dep_val=None, tiles=(100,)):
    s._touched.zero_(); s._slabs.zero_()
    NWn = 4096 // 32
    tw = s._touched.view(2, -1)                   # two parities
    slabs = s._slabs.view(-1, 4096, 32)           # [tile][ant][row]
    for tile in tiles:
        tw[3 & 1, tile * NWn:(tile + 1) * NWn] = -1
        slabs[tile] = words_fn().to(slabs.device)
    s._okbits.fill_(-1)
    if dep_val is not None:
        s._deposit.fill_(dep_val)
    run(tag)
    dump_prof(tag, 3)
    s._okbits.copy_(ok)
g = torch.Generator().manual_seed(1)
def two_random():
    a = torch.randint(0, 32, (4096, 32), generator=g); b = torch.randint(0, 32, (4096, 32), generator=g)
    return ((1 << a) | (1 << b)).to(torch.int32)
def hot_lane():
    a = torch.randint(0, 32, (4096, 32), generator=g)
    return ((1 << a) | (1 << 5)).to(torch.int32)
def all_bits():
    return torch.full((4096, 32), -1, dtype=torch.int32)
def one_row_only():
    w = torch.zeros((4096, 32), dtype=torch.int32); w[:, 3] = 1 << 5
    return w
NOFLUSH = True
print("--- no L2 flush between runs from here")
synth("one tile, 4096 ants, two random bits per word (128 rounds)", two_random, 1e-3)
synth("one tile, 4096 ants, lane 5 + a random bit", hot_lane, 1e-3)
synth("one tile, 4096 ants, all bits", all_bits, 1e-3)
synth("one tile, 4096 ants, a single row has a single bit", one_row_only, 1e-3)
synth("one tile, 4096 ants, all bits, deposits 0.0", all_bits, 0.0)
synth("16 tiles, 4096 ants, all bits", all_bits, 1e-3, tiles=tuple(range(40, 56)))
