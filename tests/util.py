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
            if p < 0:
                continue
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


class NumpyAccumulator:
    """CPU stand-in for engine.Accumulator with the same halo / compact interface (gloo tests of multigpu.py)"""

    def __init__(self, L, W, step, first_window, pair_tbl, z100, mfe_dcal, ed100):
        from scanfold_b200.foldstep import split_exact
        n = len(pair_tbl)
        self.W, self.step, self.nt0, self.n_nt, self.ncol = W, step, first_window * step, (n - 1) * step + W, 2 * W - 1
        self.count = np.zeros((self.n_nt, self.ncol), dtype=np.int32)
        self.first = np.full((self.n_nt, self.ncol), 0x7F7F7F7F, dtype=np.int32)
        self.sums = np.zeros((6, self.n_nt, self.ncol), dtype=np.int64)
        parts = [split_exact(z100), split_exact(mfe_dcal), split_exact(ed100)]
        for s in range(n):
            w = first_window + s
            for pos in range(W):
                row = s * step + pos
                p = int(pair_tbl[s][pos])
                if p < 0:
                    continue
                col = ((p - 1) - pos if p else 0) + W - 1
                self.count[row, col] += 1
                self.first[row, col] = min(self.first[row, col], w)
                for q in range(3):
                    self.sums[2 * q, row, col] += parts[q][0][s]
                    self.sums[2 * q + 1, row, col] += parts[q][1][s]

    def empty_tensors(self, n_rows):
        import torch
        n = n_rows * self.ncol
        return (torch.empty(n, dtype=torch.int32), torch.empty(n, dtype=torch.int32), torch.empty(6 * n, dtype=torch.int64))

    def export_tensors(self, row0, n_rows):
        import torch
        sl = slice(row0, row0 + n_rows)
        return (torch.from_numpy(self.count[sl].reshape(-1).copy()), torch.from_numpy(self.first[sl].reshape(-1).copy()),
                torch.from_numpy(self.sums[:, sl].reshape(-1).copy()))

    def merge_tensors(self, row0, n_rows, count, first, sums):
        sl = slice(row0, row0 + n_rows)
        self.count[sl] += count.numpy().reshape(n_rows, self.ncol)
        self.first[sl] = np.minimum(self.first[sl], first.numpy().reshape(n_rows, self.ncol))
        self.sums[:, sl] += sums.numpy().reshape(6, n_rows, self.ncol)

    def compact(self, row0=0, n_rows=None):
        if n_rows is None:
            n_rows = self.n_nt - row0
        cnt = self.count[row0:row0 + n_rows]
        rows, cols = np.nonzero(cnt)
        nparts = np.bincount(rows, minlength=n_rows)
        partner = self.nt0 + row0 + rows + 1 + cols - (self.W - 1)
        return (nparts, partner, cnt[rows, cols], self.first[row0:row0 + n_rows][rows, cols],
                self.sums[:, row0:row0 + n_rows][:, rows, cols])


def dinucl_shuffle(s, rng):
    """Altschul-Erikson dinucleotide shuffle (uniform over the Eulerian paths of the dinucleotide multigraph), the
    algorithm the reference uses for --type di (ScanFoldFunctions.py:155-277).  Test-side host implementation that
    feeds parity shuffles and host z-score samples; `rng` is a random.Random."""
    n = len(s)
    if n < 3:
        return s
    succ = {}
    for a, b in zip(s[:-1], s[1:]):
        succ.setdefault(a, []).append(b)
    last = s[-1]
    verts = sorted(set(s))
    while True:                                   # last edge per vertex: must form a tree towards `last`
        last_edge = {v: rng.choice(succ[v]) for v in verts if v != last}
        ok = True
        for v in last_edge:
            u, hops = v, 0
            while u != last and hops <= len(verts):
                u = last_edge[u]
                hops += 1
            ok &= u == last
        if ok:
            break
    lists = {}
    for v in verts:
        lst = list(succ.get(v, []))
        if v in last_edge:
            lst.remove(last_edge[v])
        rng.shuffle(lst)
        if v in last_edge:
            lst.append(last_edge[v])
        lists[v] = lst
    out, cur = [s[0]], s[0]
    ptr = {v: 0 for v in verts}
    for _ in range(n - 2):
        nxt = lists[cur][ptr[cur]]
        ptr[cur] += 1
        out.append(nxt)
        cur = nxt
    out.append(last)
    return "".join(out)
