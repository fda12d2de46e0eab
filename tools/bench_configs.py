#!/usr/bin/env python
"""Per-config throughput of the scanning hot path (SURVEY 8d "Reported numbers per config"): windows/s, folds/s and
DP cells/s for BASELINE.json configs C1..C5 on one GPU.  C2 is bench.py's own workload; this script covers the rest
on BOUNDED window ranges of the named records (the slow long-window cases would take hours in full) and prints one
JSON line per case.   usage: python tools/bench_configs.py [--max-seconds S] > profiles/rNN_configs.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scanfold_b200 import engine, foldstep, pipeline, scan, workcount  # noqa: E402


def synth(L, seed, comp=(0.25, 0.25, 0.25, 0.25)):
    rng = np.random.Generator(np.random.PCG64(seed))
    return "".join("ACGU"[k] for k in rng.choice(4, size=L, p=np.array(comp) / sum(comp)))


def reactivities(L, seed=2003):
    rng = np.random.default_rng(seed)
    v = np.clip(rng.exponential(0.4, L), 0, 4)
    v[rng.random(L) < 0.05] = -999.0
    return [-999.0] + v.tolist()


def hard_constraints(L, seed=2005):
    rng = np.random.default_rng(seed)
    return "".join("x" if u < 0.10 else "." for u in rng.random(L))


CASES = [
    # name, L, seed, composition, W, r, shuffle, extras, windows to time (None = all)
    ("C1", 1000, 1001, (1, 1, 1, 1), 120, 100, "mono", {}, None),
    ("C2", 29903, 1002, (0.299, 0.184, 0.196, 0.321), 120, 100, "mono", {}, 6000),
    ("C3", 10000, 1003, (1, 1, 1, 1), 200, 50, "di", {"react": True}, 3000),
    ("C5/W40", 100000, 1005, (1, 1, 1, 1), 40, 100, "mono", {"hc": True}, 30000),
    ("C5/W120", 100000, 1005, (1, 1, 1, 1), 120, 100, "mono", {"hc": True}, 6000),
    ("C5/W300", 100000, 1005, (1, 1, 1, 1), 300, 100, "mono", {"hc": True}, 1200),
    ("C5/W600", 100000, 1005, (1, 1, 1, 1), 600, 100, "mono", {"hc": True}, 300),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    engine.init(0)
    for name, L, seed, comp, W, r, stype, extra, nwin in CASES:
        if args.only and name not in args.only.split(","):
            continue
        seq = synth(L, seed, comp)
        total = scan.n_windows_of(L, W, 1)
        n = total if nwin is None else min(nwin, total)
        kw = dict(shuffle_type=stype, seed=42, first_window=0, n_windows=n, final_window=(n == total))
        if extra.get("react"):
            kw.update(react=reactivities(L), shape_m=0.8, shape_b=-0.2)
        if extra.get("hc"):
            kw.update(hc=hard_constraints(L))
        plan = engine.ScanPlan(seq, W, 1, r, want_pf=True, **kw)
        plan.run()                                    # warm-up (allocations, first launches)
        plan.run()
        ms_dev, ms_mfe, launches = plan.ms_total, plan.ms_mfe, plan.n_launches
        plan.close()
        t0 = time.perf_counter()                      # end to end through the host-buffer API + accumulation
        t = scan.scan_record(seq, W, 1, r, **kw)
        ptable = pipeline.partner_table_gpu(L, t)
        ms_e2e = (time.perf_counter() - t0) * 1e3
        folds = (n + (1 if n == total else 0)) * (r + 1) + (n if (extra.get("hc") or extra.get("react")) else 0)
        print(json.dumps({
            "config": name, "record_nt": L, "window": W, "shuffles": r, "shuffle_type": stype,
            "constraints": "deigan" if extra.get("react") else ("hc x 10%" if extra.get("hc") else None),
            "windows_timed": n, "windows_total": total,
            "windows_per_s": n / (ms_dev * 1e-3), "folds_per_s": folds / (ms_dev * 1e-3),
            "dp_cells_per_s": folds * workcount.cells(W) / (ms_dev * 1e-3),
            "ms_device": ms_dev, "ms_mfe_kernels": ms_mfe, "launches": launches,
            "e2e_windows_per_s": n / (ms_e2e * 1e-3), "partner_entries": int(len(ptable.partner)),
            "full_record_seconds_at_this_rate": total / (n / (ms_dev * 1e-3))}), flush=True)


if __name__ == "__main__":
    main()
