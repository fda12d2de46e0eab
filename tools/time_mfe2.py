"""times the energy-only fold path (device time of sfb_scan_plan_run without PF) for the current SFB_MFE2_TEAM"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanfold_b200 import engine
engine.init(0)
rng = np.random.default_rng(1)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 120
L = 6000 + W - 1
seq = "".join("ACGU"[k] for k in rng.choice(4, size=L, p=[0.299, 0.184, 0.196, 0.321]))
plan = engine.ScanPlan(seq, W, 1, 100, want_pf=False, final_window=False)
for rep in range(4):
    plan.run()
    print("team", os.environ.get("SFB_MFE2_TEAM", "2"), "W", W, "rep", rep, "ms_total %.1f ms_mfe %.1f  folds/s %.0f" % (
        plan.ms_total, plan.ms_mfe, 6000 * 101 / (plan.ms_mfe * 1e-3)))
plan.close()
