"""times the energy-only fold path (device time of sfb_scan_plan_run without PF).  usage: time_mfe.py W n_windows reps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanfold_b200 import engine
engine.init(0)
rng = np.random.default_rng(1)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 120
NWIN = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 4
L = NWIN + W - 1
seq = "".join("ACGU"[k] for k in rng.choice(4, size=L, p=[0.299, 0.184, 0.196, 0.321]))
plan = engine.ScanPlan(seq, W, 1, 100, want_pf=False, final_window=False)
for rep in range(REPS):
    plan.run()
    print("engine", os.environ.get("SFB_MFE_ENGINE", "3"), "W", W, "rep", rep, "ms_total %.1f ms_mfe %.1f  folds/s %.0f" % (
        plan.ms_total, plan.ms_mfe, NWIN * 101 / (plan.ms_mfe * 1e-3)))
plan.close()
