"""The ScanFold-Fold step: per-nucleotide partner statistics -> best partner -> competition -> final
partners.  Restates ScanFold.py:1036-1453 (aggregation :1051-1260, competition :1273-1453) and the helpers
competing_pairs / best_basepair (ScanFoldFunctions.py:343-474) on top of the accumulators the CUDA library
produces (sfb_accumulate_*), replacing the reference's per-record dict-of-lists churn and its O(N^2)
competition scans by O(N * partners) array code that picks exactly the same winners.

Numerical contract (what makes the outputs byte-identical to the reference run under Python >= 3.12):
  * every per-partner sum is the correctly rounded exact sum of the window values (CPython >= 3.12 `sum()`
    uses compensated summation, which is exact here); the accumulators carry each value d = k/100 as the
    exact pair A = rint(d * 2^20), B = (d - A * 2^-20) * 2^59, so the sum is order independent and a
    multi-GPU reduction is bit exact;
  * every mean is statistics.mean = the correctly rounded exact rational mean;
  * ties in `min(dict, key=dict.get)` go to the partner seen first (lowest window index).
"""
import bisect

import numpy as np

SCALE_A = 2.0 ** -20
SCALE_B = 2.0 ** -59


class PartnerTable:
    """Compact per-nucleotide partner statistics (CSR over covered nucleotides 1..n_nt, 1-based coordinates).
    Within a nucleotide the entries are in first-seen order (ascending first window index)."""

    def __init__(self, nt_ptr, partner, count, first_seen, sums, coord=None):
        self.nt_ptr = np.asarray(nt_ptr, dtype=np.int64)     # [n_nt + 1]
        self.partner = np.asarray(partner, dtype=np.int64)   # [M] partner coordinate (== own coordinate: unpaired)
        self.count = np.asarray(count, dtype=np.int64)       # [M] windows holding this pair
        self.first_seen = np.asarray(first_seen, dtype=np.int64)
        self.sums = np.asarray(sums, dtype=np.int64)         # [6, M]: zA zB mfeA mfeB edA edB
        self.n_nt = len(self.nt_ptr) - 1
        # 1-based coordinate of every row; rows are the keys of the reference's bp_dict in ascending order.  A nucleotide
        # no window left a record for (only all-N windows cover it, Appendix B Q10) has no row.
        self.coord = np.arange(1, self.n_nt + 1, dtype=np.int64) if coord is None else np.asarray(coord, dtype=np.int64)


def table_from_compact(nparts, partner, count, first_seen, sums, nt0=0):
    """sfb_accumulate_fetch arrays (entries in column order) of the nucleotides nt0, nt0+1, .. (0-based) -> PartnerTable
    in first-seen order; nucleotides without any record get no row"""
    nparts = np.asarray(nparts, dtype=np.int64)
    own = np.repeat(np.arange(len(nparts)), nparts)
    order = np.lexsort((np.asarray(first_seen), own))
    covered = nparts > 0
    ptr = np.concatenate([[0], np.cumsum(nparts[covered])])
    coord = nt0 + 1 + np.nonzero(covered)[0]
    return PartnerTable(ptr, np.asarray(partner)[order], np.asarray(count)[order], np.asarray(first_seen)[order],
                        np.asarray(sums)[:, order], coord)


def concat_tables(tables):
    """per-shard tables over consecutive nucleotide ranges -> one table"""
    ptr = [np.zeros(1, dtype=np.int64)]
    base = 0
    for t in tables:
        ptr.append(t.nt_ptr[1:] + base)
        base += t.nt_ptr[-1]
    return PartnerTable(np.concatenate(ptr), np.concatenate([t.partner for t in tables]),
                        np.concatenate([t.count for t in tables]), np.concatenate([t.first_seen for t in tables]),
                        np.concatenate([t.sums for t in tables], axis=1), np.concatenate([t.coord for t in tables]))


def split_exact(values100):
    """int k (value = k/100) -> (A, B) int64 with fl(k/100) == A * 2^-20 + B * 2^-59 exactly"""
    d = np.asarray(values100, dtype=np.float64) / 100.0
    a = np.rint(d * 2.0 ** 20)
    b = (d - a * SCALE_A) * 2.0 ** 59
    return a.astype(np.int64), b.astype(np.int64)


def _exact_sum(a, b):
    """correctly rounded value of the exact sum held as (sum A, sum B)"""
    return a.astype(np.float64) * SCALE_A + b.astype(np.float64) * SCALE_B


_THRESHOLDS = np.array([-2.0, -1.0, 0.0, 1.0, 2.0, 10.0])


def _exact_mean(a, b, n):
    """statistics.mean of the values behind (sum A, sum B): correctly rounded exact rational mean.
    Fast float64 path, exact big-integer path wherever a later rounding / comparison could see the difference."""
    s = _exact_sum(a, b)
    m = s / n
    x100 = m * 100.0
    x1e6 = m * 1e6
    risky = (np.abs(x100 - np.floor(x100) - 0.5) < 1e-6) | (np.abs(x1e6 - np.floor(x1e6) - 0.5) < 1e-3)
    for t in _THRESHOLDS:
        risky |= np.abs(m - t) < 1e-9
    for k in np.nonzero(risky)[0]:
        num = (int(a[k]) << 39) + int(b[k])
        m[k] = num / (int(n[k]) << 59)          # int / int is correctly rounded in Python
    return m


class NtResult:
    """Per-nucleotide outcome of the aggregation (index 0 <-> coordinate 1)."""
    pass


def aggregate(table, seq, log_total=None, sirna_log=None, by_ed=False):
    """ScanFold.py:1051-1260.  Returns an object with, per covered nucleotide k (arrays of length n_nt, coordinates in
    .coord):
       part        best partner coordinate (== k: unpaired)           best_coordinate
       cov_z       sum z / #total windows of the best partner         best_total_window_mean_bps[k].zscore
                   (--by_ed: sum ED / #total windows, ScanFold.py:1190-1215,1251-1255)
       mean_z      mean z over the windows holding the best pair      best_bps[k].zscore
       mean_mfe, mean_ed                                              .mfe / .ed of both dictionaries
    and writes the .ScanFold.log / .ntPairCounts.log text when file objects are given."""
    n_nt = table.n_nt
    ptr = table.nt_ptr
    M = len(table.partner)
    own = np.repeat(table.coord, np.diff(ptr))
    cnt = table.count
    sum_z = _exact_sum(table.sums[4], table.sums[5]) if by_ed else _exact_sum(table.sums[0], table.sums[1])
    want_logs = log_total is not None or sirna_log is not None
    if want_logs:      # the log prints the means of every partner entry; otherwise only the best partner's are needed
        mean_z = _exact_mean(table.sums[0], table.sums[1], cnt)
        mean_mfe = _exact_mean(table.sums[2], table.sums[3], cnt)
        mean_ed = _exact_mean(table.sums[4], table.sums[5], cnt)
    total_windows = np.add.reduceat(cnt, ptr[:-1]) if M else np.zeros(0, dtype=np.int64)
    num_bp = np.add.reduceat((table.partner != own).astype(np.int64), ptr[:-1])
    cov_z = sum_z / np.repeat(total_windows, np.diff(ptr))
    # first minimum of cov_z within each nucleotide (dict order = first-seen order)
    seg_min = np.minimum.reduceat(cov_z, ptr[:-1])
    is_min = cov_z == np.repeat(seg_min, np.diff(ptr))
    idx = np.arange(M, dtype=np.int64)
    best = np.minimum.reduceat(np.where(is_min, idx, M), ptr[:-1])

    if want_logs:
        _write_logs(table, seq, own, cnt, sum_z, mean_z, mean_mfe, mean_ed, cov_z, total_windows, num_bp,
                    log_total, sirna_log, by_ed)

    res = NtResult()
    res.n_nt = n_nt
    res.coord = table.coord
    res.part = table.partner[best]
    res.cov_z = cov_z[best]
    if want_logs:
        res.mean_z, res.mean_mfe, res.mean_ed = mean_z[best], mean_mfe[best], mean_ed[best]
    else:
        cb = cnt[best]
        res.mean_z = _exact_mean(table.sums[0][best], table.sums[1][best], cb)
        res.mean_mfe = _exact_mean(table.sums[2][best], table.sums[3][best], cb)
        res.mean_ed = _exact_mean(table.sums[4][best], table.sums[5][best], cb)
    res.total_windows = np.asarray(total_windows, dtype=np.int64)
    res.num_bp = np.asarray(num_bp, dtype=np.int64)
    return res


def _r2(x):
    return str(round(float(x), 2))


def _r2_list(values):
    """[str(round(float(x), 2)) for x in values], vectorised: the two-decimal value is an integer number of hundredths
    (a few thousand distinct ones per record), formatted once each by Python itself; anything close to a rounding
    boundary, not finite or large goes through round() directly"""
    v = np.asarray(values, dtype=np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        x = v * 100.0
        k = np.rint(x)
        slow = ~np.isfinite(v) | (np.abs(v) > 1e9) | (np.abs(x - np.floor(x) - 0.5) < 1e-6)
    ki = np.where(slow, 0, k).astype(np.int64)
    neg0 = (ki == 0) & (np.signbit(v) | np.signbit(k))      # round(-0.001, 2) is -0.0 and prints "-0.0"
    uniq, inv = np.unique(ki, return_inverse=True)
    table = np.array([str(u / 100.0) for u in uniq.tolist()], dtype=object)
    out = table[inv]
    out[neg0] = "-0.0"
    for m in np.nonzero(slow)[0]:
        out[m] = str(round(float(v[m]), 2))
    return out.tolist()


def _write_logs(table, seq, own, cnt, sum_z, mean_z, mean_mfe, mean_ed, cov_z, total_windows, num_bp, log_total,
                sirna_log, by_ed=False):
    """log lines of ScanFold.py:1150-1184; --by_ed: :1155-1159,1190-1212 (sum_z / cov_z then hold the ED sums; the column
    order avgMFE, avgZ, avgED is the same in both modes).  One string per partner entry is assembled column-wise (the
    record's log has a line per entry: 10^6 for a 30-kb genome), then cut per nucleotide."""
    ptr = table.nt_ptr.tolist()
    partner = table.partner
    head = "SumED\tSumED/#TotalWindows" if by_ed else "SumZ\tSumZ/#TotalWindows"
    coord = table.coord.tolist()
    sq = np.frombuffer(seq.encode(), dtype="S1")
    own_s = list(map(str, np.asarray(own).tolist()))
    part_l = partner.tolist()
    bp_s = ["NoBP" if j == k else str(j) for j, k in zip(part_l, np.asarray(own).tolist())]
    nuc_j = sq[partner - 1].astype("U1").tolist()
    cols = (own_s, bp_s, nuc_j, list(map(str, np.asarray(cnt).tolist())), _r2_list(mean_mfe), _r2_list(mean_z),
            _r2_list(mean_ed), _r2_list(sum_z), _r2_list(cov_z))
    lines = list(map("\t".join, zip(*cols)))
    out = []
    sirna = []
    tw = np.asarray(total_windows).tolist()
    nb = np.asarray(num_bp).tolist()
    for k0 in range(table.n_nt):
        k = coord[k0]
        nuc = seq[k - 1]
        out.append("\ni-nuc\tBP(j)\tNuc\t#BP_Win\tavgMFE\tavgZ\tavgED\t%s\tBPs= %d\n" % (head, nb[k0]))
        out.append("nt-%d\t-\t%s\t%d\t-\t-\t-\t-\t-\n" % (k, nuc, tw[k0]))
        sirna.append("%d\t%s\t%d\t%d\n" % (k, nuc, tw[k0], nb[k0]))
        if ptr[k0 + 1] > ptr[k0]:
            out.append("\n".join(lines[ptr[k0]:ptr[k0 + 1]]))
            out.append("\n")
    if log_total is not None:
        log_total.write("".join(out))
    if sirna_log is not None:
        sirna_log.write("".join(sirna))


FINAL_PARTNERS_HEADER = ("i\tbp(i)\tbp(j)\tavgMFE\tavgZ\tavgED\t*Indicates most favorable bp has competition; bp(j) has "
                         "more favorable partner or is more likely to be unpaired\n")


class FinalPartners:
    """final_partners of ScanFold.py:1404-1442, one entry per covered nucleotide k (index k-1):
       i, j     icoordinate / jcoordinate of the stored NucPair (i == j: unpaired)
       z, mfe, ed   its metrics
    """
    pass


def compete(res, seq, log_win=None):
    """ScanFold.py:1273-1442 with competition == 1.  `res` comes from aggregate()."""
    n = res.n_nt
    coord = res.coord
    part = res.part
    z = res.cov_z
    top = int(max(coord.max(), part.max())) + 2 if n else 2
    pos = np.full(top, -1, dtype=np.int64)          # coordinate -> row
    pos[coord] = np.arange(n)
    # inverse index: inv[c] = ascending nucleotides whose best partner is c   (competing_pairs scans)
    order = np.argsort(part, kind="stable")
    sorted_part = part[order]
    keys_of = coord[order]

    def comp(c):
        """entries of best_total_window_mean_bps touching coordinate c, in dictionary (ascending) order
        (competing_pairs, ScanFoldFunctions.py:343-357)"""
        lo, hi = np.searchsorted(sorted_part, (c, c + 1))
        lst = keys_of[lo:hi].tolist()
        r = pos[c] if c < top else -1
        if r >= 0 and part[r] != c:         # entry c itself (icoordinate == c) unless already listed
            bisect.insort(lst, c)
        return lst

    comp_cache = {}

    def comp_c(c):
        r = comp_cache.get(c)
        if r is None:
            r = comp_cache[c] = comp(c)
        return r

    fin = FinalPartners()
    fin.coord = coord
    fin.i = np.zeros(n, dtype=np.int64)
    fin.j = np.zeros(n, dtype=np.int64)
    fin.z = np.zeros(n)
    fin.mfe = np.zeros(n)
    fin.ed = np.zeros(n)
    zl = z.tolist()
    partl = part.tolist()
    posl = pos.tolist()
    coordl = coord.tolist()
    lines = []
    if log_win is not None:
        lines.append(FINAL_PARTNERS_HEADER)
    for r0 in range(n):
        k = coordl[r0]
        i, j = k, partl[r0]
        best_m, best_z = -1, None
        for c in (i, j):
            for m in comp_c(c):
                for cc in (partl[posl[m]], m):
                    for mm in comp_c(cc):
                        zz = zl[posl[mm]]
                        if best_z is None or zz < best_z:
                            best_m, best_z = mm, zz
        wi, wj = best_m, partl[posl[best_m]]
        if k != wi and k != wj:
            if log_win is not None:
                lines.append("nt-%d*:\t%d\t%d\t%s\t%s\t%s\n" % (k, i, j, _r2(res.mean_mfe[r0]), _r2(res.mean_z[r0]),
                                                             _r2(res.mean_ed[r0])))
            fin.i[r0] = fin.j[r0] = k
        else:
            if log_win is not None:
                lines.append("nt-%d:\t%d\t%d\t%s\t%s\t%s\n" % (k, wi, wj, _r2(res.mean_mfe[r0]), _r2(res.mean_z[r0]),
                                                            _r2(res.mean_ed[r0])))
            fin.i[r0], fin.j[r0] = wi, wj
        w = posl[wi]
        fin.z[r0] = res.mean_z[w]
        fin.mfe[r0] = res.mean_mfe[w]
        fin.ed[r0] = res.mean_ed[w]
    if log_win is not None:
        log_win.write("".join(lines))
    return fin
