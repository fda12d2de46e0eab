"""device time of the partition-function kernel inside a scan plan (ms_total - ms_mfe with 2 shuffles per window).
usage: time_pf.py W n_windows"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanfold_b200 import engine
engine.init(0)
rng = np.random.default_rng(1)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 120
NWIN = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
L = NWIN + W - 1
seq = "".join("ACGU"[k] for k in rng.choice(4, size=L, p=[0.299, 0.184, 0.196, 0.321]))
plan = engine.ScanPlan(seq, W, 1, 2, want_pf=True, final_window=False)
for rep in range(4):
    plan.run()
    pf_ms = plan.ms_total - plan.ms_mfe
    print("pf engine", os.environ.get("SFB_PF_ENGINE", "2"), "W", W, "rep", rep, "pf+misc %.2f ms  (%.0f windows/s)" % (pf_ms, NWIN / (pf_ms * 1e-3)))
res = plan.fetch()
print("ed[0]=%.6f dG[0]=%.6f" % (res.ed[0], res.ensemble_dG[0]))
plan.close()
