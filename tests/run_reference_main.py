"""Runs the UNMODIFIED reference main.py (path in argv[1]) with `maaco_path_planing_b200/dropin` on sys.path in
place of the reference's own modules, and dumps what it computed as JSON (argv[2]).  Test helper."""
import json
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maaco_path_planing_b200", "dropin"))
ns = runpy.run_path(sys.argv[1], run_name="__main__")


def clean(v):
    if isinstance(v, (list, tuple)):
        return [clean(x) for x in v]
    if hasattr(v, "tolist"):
        return clean(v.tolist())
    if isinstance(v, float):
        return v if v == v and abs(v) != float("inf") else repr(v)
    return v


keep = {}
for k, v in ns.items():
    if any(k.startswith(p) for p in ("maaco_", "mpa_", "astar_", "dijkstra_", "ga_", "pso_")) and \
            any(t in k for t in ("_path_", "_len_", "_turns_", "_sp_", "_dp_", "_fit_")):
        keep[k] = clean(v)
for k in ("current_test_grid_fig7_processed", "grid_map_fig13_processed"):
    keep[k] = clean(ns[k])
for k in ("maaco_solver_fig7", "maaco_solver_fig13", "mpa_solver_fig7", "ga_solver_fig7", "pso_solver_fig7"):
    s = ns[k]
    keep[k + ".curve"] = clean(list(getattr(s, "convergence_curve_data", None) or getattr(s, "convergence_curve", [])))
json.dump(keep, open(sys.argv[2], "w"))
