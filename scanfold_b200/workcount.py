"""Algorithmic work of one MFE fold: DP cells and min-plus relaxations.

A relaxation is one candidate term acc = min(acc, a + b) of the Zuker recursions: one hairpin term per
cell, one per admissible interior-loop candidate (u1 + u2 <= 30, q - p > 3), one per FML split, one per
F5 term.  `dense` counts every candidate regardless of pairability (a function of W only); `useful`
counts only candidates whose operands can be finite (canonical pairs; FML entries that contain a pair).
bench.py divides these by the measured kernel time; nothing here touches the GPU or the oracle.
"""
import numpy as np

TURN = 3
MAXLOOP = 30
_PAIR = np.zeros((5, 5), dtype=bool)
for _a, _b in ((2, 3), (3, 2), (3, 4), (4, 3), (1, 4), (4, 1)):  # CG GC GU UG AU UA with A=1 C=2 G=3 U=4
    _PAIR[_a, _b] = True


def cells(W):
    """(i, j) with j - i > TURN"""
    return max(0, (W - 4) * (W - 3) // 2)


def dense_relaxations(W):
    """R(W) of SURVEY Appendix C: 72,508 / 2,475,988 / 8,783,468 / 22,937,818 / 116,800,868 for 40..600"""
    total = 0
    for d in range(TURN + 1, W):
        ncell = W - d
        inter = 0
        for u1 in range(0, min(MAXLOOP, d - 2 - TURN - 1) + 1):
            # q - p = d - 2 - u1 - u2 > TURN  ->  u2 <= d - 3 - TURN - u1
            u2max = min(MAXLOOP - u1, d - 3 - TURN - u1)
            if u2max >= 0:
                inter += u2max + 1
        splits = max(0, d - 2 * TURN - 2)
        total += ncell * (1 + inter + splits + 1)  # hairpin + interior + FML splits + F5 term
    return total


def encode(seq):
    tab = np.zeros(256, dtype=np.uint8)
    for ch, v in (("A", 1), ("C", 2), ("G", 3), ("U", 4), ("T", 4)):
        tab[ord(ch)] = v
        tab[ord(ch.lower())] = v
    a = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), dtype=np.uint8)
    return tab[a]


def useful_relaxations(seq):
    """Relaxations of one unconstrained fold whose operands are finite (what the oracle's counter calls useful)."""
    S = encode(seq)
    W = len(S)
    ii, jj = np.meshgrid(np.arange(W), np.arange(W), indexing="ij")
    P = _PAIR[S[:, None], S[None, :]] & (jj - ii > TURN)
    useful = int(P.sum())  # hairpin term of every pairable cell
    # interior candidates: pairable outer (i,j) x pairable inner (i+1+u1, j-1-u2)
    for u1 in range(MAXLOOP + 1):
        for u2 in range(MAXLOOP + 1 - u1):
            a, b = 1 + u1, 1 + u2
            if a + b >= W:
                continue
            useful += int((P[:W - a, b:] & P[a:, :W - b]).sum())
    # FML[i,j] is finite iff [i,j] contains a pairable pair
    A = P.copy()
    for d in range(TURN + 1, W):
        i = np.arange(W - d)
        A[i, i + d] |= A[i + 1, i + d] | A[i, i + d - 1]
    for d in range(2 * TURN + 3, W):
        i = np.arange(W - d)
        for k in range(TURN + 1, d - 1 - TURN):
            useful += int((A[i, i + k] & A[i + k + 1, i + d]).sum())
    useful += int(P.sum())  # F5 terms with a finite C[i,j]
    return useful
