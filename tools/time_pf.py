"""times the partition-function kernel alone.  usage: time_pf.py W n_windows"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanfold_b200 import engine
sys.path.insert(0, "tests")
from util import rand_seqs
engine.init(0)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 120
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
seqs = rand_seqs(11, N, W)
for rep in range(3):
    t0 = time.time(); r = engine.pf_batch(seqs); t1 = time.time()
    print("engine", os.environ.get("SFB_PF_ENGINE", "2"), "W", W, "n", N, "%.1f ms  (%.0f windows/s incl. copies)  ed[0]=%.6f dG[0]=%.6f" % (
        (t1 - t0) * 1e3, N / (t1 - t0), r["ed"][0], r["dG"][0]))
