import random

import numpy as np


def rand_seq(rng, n, alphabet="ACGU"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def rand_seqs(seed, n_seq, n, gc_rich=False):
    rng = random.Random(seed)
    out = []
    for k in range(n_seq):
        alpha = "ACGU"
        if gc_rich and k % 3 == 0:
            alpha = "GGCCAU"
        out.append(rand_seq(rng, n, alpha))
    return out


def db_from_pt(pt):
    idx = np.arange(1, len(pt) + 1)
    out = np.full(len(pt), ord("."), dtype=np.uint8)
    out[pt > idx] = ord("(")
    out[(pt > 0) & (pt < idx)] = ord(")")
    return out.tobytes().decode()
