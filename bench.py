#!/usr/bin/env python
"""bench.py -- path evaluations / second of the B200 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload maaco] [--impl reference]

Default workload (N=1): BASELINE config 4 -- MAACO, 4096 ants on a 512x512 synthetic block map
(`blocks(512, 0.20, seed=4000)`), parameters of main.py:34-38.  One *step* = one colony pass
(tour construction for every ant + best tracking + pheromone update); one path evaluation = one
`_construct_ant_solution_maaco` call incl. its share of the pheromone update.
N>1: one process per GPU (torchrun), the colony is sharded with a fixed 4096 ants per GPU (weak
scaling; `--strong` keeps 4096 ants in total) and exchanges visited-bitmap slices + tau slices per pass.

`--impl reference` times the CPU oracle port (oracle/mpp_oracle.c, OpenMP over ants, all host
threads) on a bounded sample of the same workload -- the reference itself is pure Python and
cannot travel to the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAACO_PARAMS = dict(alpha=1.0, beta=7.0, rho=0.1, Q=2.5, a_turn_coef=1.0, wh_max=0.9, wh_min=0.2,
                    k_h_adaptive=0.9, q0_initial=0.5, C0_initial_pheromone=0.1)   # main.py:34-38
BYTES_PER_ANT_STEP = 148          # 8x8 B tau + 8x8 B E-table + 2x8 B uniforms + 4 B path cell
METRIC = "path evals/sec (MAACO ant tours, MPA/PSO/GA fitness) at 1/2/4/8 B200 vs CPU"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_maaco(grid, n_ants, seconds_budget, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    orc = O.MaacoOracle(grid, n_ants, 100, seed=4, threads=threads, **MAACO_PARAMS)
    orc.iterate(1)                                    # warm-up pass (page faults, OpenMP pool)
    t0 = time.perf_counter()
    passes, it = 0, 2
    steps0 = orc.total_steps
    while True:
        orc.iterate(it)
        it += 1
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= seconds_budget or passes >= 64:
            break
    return {"evals": passes * n_ants, "seconds": dt, "passes": passes, "ant_steps": orc.total_steps - steps0}


def run_reference(args):
    """`--impl reference`: only rank 0 works; the others exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from maaco_path_planing_b200.gridmap import blocks_map
    cores = os.cpu_count() or 1
    grid = blocks_map(args.size, 0.20, seed=4000)
    per_step = max(1.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_maaco(grid, args.ants, 0.0, 0)
    tot_e, tot_s = 0, 0.0
    for _ in range(args.steps):
        r = cpu_maaco(grid, args.ants, per_step, 0)
        tot_e += r["evals"]
        tot_s += r["seconds"]
    v = tot_e / tot_s
    sample = f"{tot_e // args.ants} colony passes of {args.ants} ants on blocks({args.size},0.20,4000), OpenMP over ants"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "path evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": "path evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "path evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    total = args.ants if args.strong else args.ants * world
    return {"workload": f"MAACO colony pass, {total} ants ({total // world}/GPU) on {args.size}x{args.size} "
                        f"blocks(n,0.20,seed=4000), params main.py:34-38 (BASELINE config 4)",
            "ants_total": total, "grid": [args.size, args.size], "l2": "flushed (256 MiB write) between timed steps",
            "parallelism": f"colony sharded over {world} GPU(s)" + ("" if world == 1 else
                                                                  "; all-to-all of visited-bitmap slices + all-gather of tau slices per pass")}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from maaco_path_planing_b200 import MAACO, blocks_map

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    dev = torch.device("cuda", local)
    total_ants = args.ants if args.strong else args.ants * world
    grid = blocks_map(args.size, 0.20, seed=4000)
    K, W = args.steps, args.warmup
    solver = MAACO(grid, total_ants, K + W + 16, rng_seed=4, device=local, group=group, verbose=False, **MAACO_PARAMS)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(group=group)
            torch.cuda.synchronize(dev)

    it = 0
    for _ in range(W):
        it += 1
        solver.run_iteration(it)
    sync_all()
    steps_before = solver.total_steps()
    launches_before = solver.kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
    t_wall0 = time.time()
    sync_all()
    for k in range(K):
        flush.fill_(k & 0xff)                                   # evict L2 between timed steps (untimed)
        it += 1
        solver._enqueue_iteration(it, events=ev[k])             # events: start / after tours / before pheromone / end / after ranking
    sync_all()
    t_wall1 = time.time()
    step_ms = [e[0].elapsed_time(e[3]) for e in ev]
    tour_ms = [e[4].elapsed_time(e[1]) for e in ev]             # the tour kernel alone
    rank_ms = [e[0].elapsed_time(e[4]) for e in ev]             # the move-ranking kernel before it
    pher_ms = [e[2].elapsed_time(e[3]) for e in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX, group=group)
    total_ms = float(total_ms.item())
    ant_steps_local = solver.total_steps() - steps_before
    launches_timed = solver.kernel_launches - launches_before      # kernels of this library inside the timed region
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- e2e: the colony pass called with HOST buffers (pinned tau in, results + tau out) ----------
    n = args.size * args.size
    tau_host = torch.empty(n, dtype=torch.float64).pin_memory()
    tau_host.copy_(solver._tau[:n].cpu())
    res_host = torch.empty((total_ants, 2), dtype=torch.int64).pin_memory()
    best_host = torch.empty(8192, dtype=torch.int32).pin_memory()
    state_host = torch.empty(40, dtype=torch.uint8).pin_memory()
    K2 = max(3, min(K, 10))
    sync_all()
    t0 = time.perf_counter()
    for _ in range(K2):
        it += 1
        solver._tau[:n].copy_(tau_host, non_blocking=True)              # H2D: this pass's pheromone field
        solver.run_iteration(it)
        tau_host.copy_(solver._tau[:n], non_blocking=True)              # D2H: updated field
        res_host.copy_(solver._result, non_blocking=True)               # D2H: per-ant (length, n_cells, turns)
        best_host.copy_(solver._best_cells[:8192], non_blocking=True)   # D2H: best path so far
        state_host.copy_(solver._state, non_blocking=True)
        torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier(group=group)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX, group=group)
    e2e_s = float(e2e_s.item())
    h2d = n * 8
    d2h = n * 8 + total_ants * 16 + 8192 * 4 + 40

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if world == 1 and args.ants == 4096 and args.size == 512 and os.path.exists(tp):
        traffic = json.load(open(tp))["mpp_maaco_tour1_kernel"]["dram_bytes_per_launch"]   # from the committed ncu capture
    value = total_ants * K / (total_ms / 1e3)
    tour_avg_ms = sum(tour_ms) / K
    achieved = BYTES_PER_ANT_STEP * (ant_steps_local / K) / (tour_avg_ms / 1e3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": "path evals/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": total_ants * K2 / e2e_s, "unit": "path evals/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": K2,
                "note": "host pheromone field in, per-ant results + best path + updated field out, every pass"},
        "gpu_launches": launches_timed,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "mpp_maaco_tour1_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_unit": BYTES_PER_ANT_STEP, "units_per_launch": ant_steps_local / K,
                     "kernel_ms": tour_avg_ms,
                     "rank_kernel_ms": sum(rank_ms) / K,
                     "pheromone_kernel": {"ms": sum(pher_ms) / K,
                                          "achieved": (16.0 * n + 4.0 * solver.n_words * total_ants / world
                                                       + 16.0 * total_ants) / (sum(pher_ms) / K / 1e3) / 1e9,
                                          "unit": "GB/s (tau RMW + visited words streamed)"}},
        "ant_steps_per_s": ant_steps_local * world / (total_ms / 1e3),
    }
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        r = cpu_maaco(grid, args.ants, 12.0, 0)
        out["cpu_baseline"] = {"value": r["evals"] / r["seconds"], "unit": "path evals/s", "cores": cores, "kind": "port",
                               "sample": f"{r['passes']} colony passes of {args.ants} ants (same map/params), C oracle, "
                                         f"OpenMP over ants, {r['seconds']:.1f} s",
                               "ant_steps_per_s": r["ant_steps"] / r["seconds"]}
        r1 = cpu_maaco(grid, args.ants, 4.0, 1)
        out["cpu_baseline"]["one_core_value"] = r1["evals"] / r1["seconds"]
    if world == 1 and not args.no_extra:
        del solver, flush
        torch.cuda.empty_cache()
        out["extra"] = side_workloads(args, dev, peak)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def side_workloads(args, dev, peak):
    """The other components of the metric (reported under "extra"; the headline stays the MAACO colony):
    PSO/GA A*-connector fitness at BASELINE config 3 (512x512, N=4096, W=5, policy main.py:21-24) and one
    MPA iteration at config 2 (100x100, N=1024, params main.py:44-52)."""
    import numpy as np
    import torch
    from maaco_path_planing_b200 import GridMap, blocks_map
    from maaco_path_planing_b200.engine import SearchEngine, make_policy
    from maaco_path_planing_b200.mpa import MPA
    out = {}
    # ---- config 3: fitness evaluation of a population of random free-cell waypoint chromosomes ----
    size, N, W = args.fit_size, args.fit_pop, 5
    grid = blocks_map(size, 0.20, seed=3000 + size)
    rng = np.random.default_rng(3)
    free = np.flatnonzero(grid.ravel() != 1)
    eng = SearchEngine(GridMap(grid))
    pol = make_policy(0.3, 0.8, 1.8, 100.0)
    wps = [torch.as_tensor(free[rng.integers(0, len(free), (N, W))].astype(np.int32), device=dev) for _ in range(3)]
    eng.waypoint_fitness(wps[0], pol)
    torch.cuda.synchronize()
    eng.counters.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for k in range(1, 3):
        e0.record()
        cells, ncell, stats = eng.waypoint_fitness(wps[k], pol)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    exp, rel = eng.expansions()
    t = sum(ms) / 1e3
    ach = (72.0 * exp + 25.0 * rel) / t / 1e9
    out["pso_ga_fitness"] = {"workload": f"A*-connector fitness, {N} individuals x {W} waypoints on {size}x{size} blocks map "
                             "(BASELINE config 3)", "value": 2 * N / t, "unit": "path evals/s", "ms_per_population": 1e3 * t / 2,
                             "astar_expansions_per_s": exp / t, "expansions_per_eval": exp / (2 * N),
                             "valid_fraction": float((ncell > 0).float().mean()),
                             "roofline": {"bound": "hbm", "kernel": "mpp_waypoint_fitness_kernel", "achieved": ach, "peak": peak,
                                          "unit": "GB/s", "frac": ach / peak,
                                          "algorithmic_bytes_per_unit": "72 B/expansion + 25 B/relaxation"}}
    del eng
    # ---- config 2: MPA iterations ----
    grid = blocks_map(100, 0.20, seed=2000)
    mpa = MPA(grid, num_predators=1024, num_iterations=12, FADs_rate=0.2, P_const=0.5, levy_beta=2.0,
              turn_penalty_factor=0.1, safety_penalty_factor=0.8, min_safe_distance=1.8, diagonal_obstacle_penalty=100.0,
              rng_seed=2, verbose=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mpa.solve_path_planning()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["mpa"] = {"workload": "MPA 1024 predators x 12 iterations on 100x100 blocks map (BASELINE config 2), "
                  "sorts + best cascade on the host included", "value": mpa.predator_evaluations / dt,
                  "unit": "predator-iterations/s", "seconds": dt, "best_fitness": mpa.best_fitness_overall}
    # ---- config 5 (one GPU's share): independent 256x256 maps x 1024 ants, colonies overlapped on streams ----
    from maaco_path_planing_b200.batch import solve_maaco_batch
    n_maps, ants, iters = args.batch_maps, 1024, 5
    grids = [blocks_map(256, 0.20, seed=5000 + i) for i in range(n_maps)]
    solve_maaco_batch(grids[:4], ants, 2, MAACO_PARAMS, seeds=list(range(4)), concurrent=4)      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = solve_maaco_batch(grids, ants, iters, MAACO_PARAMS, seeds=list(range(n_maps)), concurrent=args.batch_concurrent)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["batched_maps"] = {"workload": f"{n_maps} independent 256x256 maps x {ants} ants x {iters} MAACO iterations, "
                           f"{args.batch_concurrent} colonies overlapped on CUDA streams (BASELINE config 5, one GPU's share; "
                           "includes per-map table build + H2D)", "value": n_maps * ants * iters / dt, "unit": "path evals/s",
                           "maps_per_s": n_maps / dt, "seconds": dt,
                           "solved_maps": sum(1 for r in res if r[1])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="maaco", choices=["maaco"])
    ap.add_argument("--ants", type=int, default=4096, help="ants per GPU (total with --strong)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--strong", action="store_true", help="fixed 4096-ant colony sharded over the GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the PSO/GA fitness and MPA side workloads")
    ap.add_argument("--fit-size", type=int, default=512)
    ap.add_argument("--fit-pop", type=int, default=4096)
    ap.add_argument("--batch-maps", type=int, default=48)
    ap.add_argument("--batch-concurrent", type=int, default=12)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
