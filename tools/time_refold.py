"""wall time of one whole-sequence fold (sfb_fold_long) by length.  usage: time_refold.py [L ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanfold_b200 import engine
engine.init(0)
for L in [int(a) for a in sys.argv[1:]] or [2000, 10000, 29903]:
    rng = np.random.Generator(np.random.PCG64(1002))
    seq = "".join("ACGU"[k] for k in rng.choice(4, size=L, p=[0.299, 0.184, 0.196, 0.321]))
    for rep in range(2):
        t0 = time.perf_counter()
        e, pt = engine.fold_long(seq)
        dt = time.perf_counter() - t0
    print("L %d  MFE %.2f kcal/mol  pairs %d  wall %.3f s  (%.2f G split add-min/s)" % (
        L, e / 100.0, int((pt > 0).sum()) // 2, dt, L ** 3 / 6 / dt / 1e9), flush=True)
