"""GPU check of the energy-only kernel selected by SFB_MFE_ENGINE against the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanfold_b200 import engine
sys.path.insert(0, "tests")
from util import rand_seqs
from oracle import oracle as O
engine.init(0)
Ws = [int(a) for a in sys.argv[1:]] or [16, 17, 20, 24, 31, 40, 64, 65, 80, 100, 119, 120]
for W in Ws:
    seqs = rand_seqs(7000 + W, 512, W, gc_rich=(W % 2 == 1))
    try:
        e2, _ = engine.fold_batch(seqs, structure=False)
    except Exception as ex:
        print("W", W, "energy-only failed:", ex)
        continue
    e1 = np.asarray(O.fold_batch(np.frombuffer("".join(seqs).encode(), dtype=np.uint8).reshape(len(seqs), W), n_threads=16))
    bad = np.nonzero(e1 != e2)[0]
    print("W", W, "mismatch", len(bad), "of", len(seqs), "redo", int((e2 == 0x7fffff00).sum()))
    for k in bad[:4]:
        print("  ", k, seqs[k], int(e2[k]), int(e1[k]))
W = 120
seqs = np.frombuffer("".join(rand_seqs(1, 60000, W)).encode(), dtype=np.uint8).reshape(-1, W)
for rep in range(3):
    try:
        t0 = time.time(); e2, _ = engine.fold_batch(seqs, structure=False); t1 = time.time()
    except Exception as ex:
        print("failed", ex); break
    print("energy-only 60000 folds: %.3f s  (%.0f folds/s) redo=%d" % (t1 - t0, 60000 / (t1 - t0), int((e2 == 0x7fffff00).sum())))
