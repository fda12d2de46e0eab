"""host-side breakdown of one end-to-end step (bench.py's e2e arm) on C2.  usage: time_e2e.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from scanfold_b200 import engine, foldstep, pipeline, scan, stats
engine.init(0)
seq, W, step, r, stype = bench.synth_record("C2")
L = len(seq)
def T(label, t0):
    t1 = time.perf_counter(); print("%-34s %8.1f ms" % (label, (t1 - t0) * 1e3)); return t1
for rep in range(2):
    print("--- rep", rep)
    t0 = time.perf_counter()
    plan = engine.ScanPlan(seq, W, step, r, shuffle_type=stype, seed=42, want_pf=True); t0 = T("ScanPlan create (H2D)", t0)
    plan.run(); t0 = T("plan.run (device %.1f ms)" % plan.ms_total, t0)
    res = plan.fetch(); t0 = T("plan.fetch (D2H)", t0)
    plan.close(); t0 = T("plan.close", t0)
    z, p = stats.zscore_pvalue(res.native_unconstrained_dcal, res.shuffle_dcal); t0 = T("stats.zscore_pvalue", t0)
    t = scan.table_from_result(res, 0, step, res.n - 1, True); t0 = T("table_from_result (incl. stats again)", t0)
    z100, mfe100, ed100 = pipeline.fold_inputs(t); t0 = T("fold_inputs", t0)
    acc = engine.Accumulator(L, W, step, 0, t.pair_tbl, z100, mfe100, ed100); t0 = T("Accumulator (H2D + kernel)", t0)
    comp = acc.compact(0, None); t0 = T("acc.compact (kernels + D2H)", t0)
    acc.close(); t0 = T("acc.close", t0)
    pt = foldstep.table_from_compact(*comp); t0 = T("table_from_compact", t0)
