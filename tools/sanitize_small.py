"""tiny workload for compute-sanitizer (memcheck / racecheck): a few folds through every kernel configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import rand_seqs
from scanfold_b200 import engine
engine.init(0)
for W in [int(a) for a in sys.argv[1:]] or [40, 120, 200, 250]:
    seqs = rand_seqs(5 + W, 3, W, gc_rich=True)
    e, _ = engine.fold_batch(seqs, structure=False)
    e2, pt = engine.fold_batch(seqs, structure=True)
    hc = ["".join("x" if (k * 7 + i) % 11 == 0 else "." for i in range(W)) for k in range(len(seqs))]
    e3, pt3 = engine.fold_batch(seqs, hc=hc, structure=True)
    print("W", W, e.tolist(), e2.tolist(), e3.tolist())
    if W <= 200:
        r = engine.pf_batch(seqs[:2])
        r2 = engine.pf_batch(seqs[:2], hc=hc[:2])
        print("   pf", r["ed"].tolist(), r2["ed"].tolist())
