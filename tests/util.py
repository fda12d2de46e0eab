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


def accumulate_numpy(L, W, step, first_window, pair_tbl, z100, mfe_dcal, ed100):
    """test-side reference of sfb_accumulate_begin + compact: the same compact arrays, computed with plain loops"""
    from scanfold_b200.foldstep import split_exact
    n = len(pair_tbl)
    nt0 = first_window * step
    n_nt = (n - 1) * step + W
    ncol = 2 * W - 1
    count = np.zeros((n_nt, ncol), dtype=np.int64)
    first = np.full((n_nt, ncol), 0x7F7F7F7F, dtype=np.int64)
    sums = np.zeros((6, n_nt, ncol), dtype=np.int64)
    parts = [split_exact(z100), split_exact(mfe_dcal), split_exact(ed100)]
    for s in range(n):
        w = first_window + s
        for pos in range(W):
            row = w * step + pos - nt0
            p = int(pair_tbl[s][pos])
            col = ((p - 1) - pos if p else 0) + W - 1
            count[row, col] += 1
            first[row, col] = min(first[row, col], w)
            for q in range(3):
                sums[2 * q, row, col] += parts[q][0][s]
                sums[2 * q + 1, row, col] += parts[q][1][s]
    rows, cols = np.nonzero(count)
    nparts = np.bincount(rows, minlength=n_nt)
    partner = nt0 + rows + 1 + cols - (W - 1)
    return nparts, partner, count[rows, cols], first[rows, cols], sums[:, rows, cols]
