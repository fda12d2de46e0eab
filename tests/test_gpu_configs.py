"""GPU parity at the real geometries of BASELINE.json's configs (bounded window slices, the exact generators of
tools/bench_configs.py) and the statistical equivalence of the device (Philox) shuffles with the reference's.

  C3   W=200, dinucleotide shuffles, Deigan pseudo-energies      mfe3<200> with sc + traceback, pf.cu at 200 nt
  C5   W=300 / 600 with an 'x' constraint line                   mfe3<300,FMG> masks, the long-window kernel, PF with hc
  PF   unconstrained partition function above 200 nt
  C4   a > 65,536-window record at W=120, r=100                  chunk boundary of sfb_scan_plan_run
Bar: energies / structures bit-exact, PF and ED within 1e-6 relative (north_star).  Shuffle statistics: two-sample
tests against host shuffles (random.sample / Altschul-Erikson) and against a frequency table drawn from the
reference's own dinuclShuffle (tests/golden/dishuffle_counts.json), all with fixed seeds, failing at p < 1e-3.
"""
import json
import os
import random
import sys

import numpy as np
import pytest

from util import db_from_pt, dinucl_shuffle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from bench_configs import hard_constraints, reactivities, synth  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, tol=1e-6):
    return abs(a - b) <= tol * max(1.0, abs(b))


def _oracle_energies(oracle, seqs):
    """energy-only oracle folds through the batch path (the tuned path is checked equal to the simple one on CPU)"""
    a = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).reshape(len(seqs), len(seqs[0]))
    return oracle.fold_batch(a, n_threads=os.cpu_count() or 1, fast=True)


def test_c3_w200_di_deigan(engine, oracle):
    """C3: 10-kb record, W=200, 50 dinucleotide shuffles, Deigan m=0.8 b=-0.2 -- 24 windows from the middle"""
    L, W, r, w0, n = 10000, 200, 50, 4321, 24
    seq = synth(L, 1003)
    react = np.array(reactivities(L))
    res = engine.scan(seq, W, 1, r, shuffle_type="di", seed=42, react=react, shape_m=0.8, shape_b=-0.2,
                      first_window=w0, n_windows=n, final_window=False, keep_shuffles=True)
    es = oracle.deigan(react, 0.8, -0.2)
    shuf = ["".join(map(chr, res.shuffles[w, k])) for w in range(n) for k in range(r)]
    e_sh = _oracle_energies(oracle, shuf).reshape(n, r)
    assert np.array_equal(e_sh, res.shuffle_dcal)
    for w in range(n):
        frag = seq[w0 + w:w0 + w + W]
        for k in range(0, r, 7):
            sh = shuf[w * r + k]
            assert sh[0] == frag[0] and sh[-1] == frag[-1] and sorted(sh) == sorted(frag)
        scw = np.zeros(W + 1, dtype=np.int32)
        for p in range(1, W):                       # Q7: +1 shift, position W reads past the slice
            g = w0 + w + 1 + p
            scw[p] = es[g] if g <= L else 0
        eo, so = oracle.mfe(frag, sc_stack=scw)
        assert res.mfe_dcal[w] == eo and db_from_pt(res.pair_tbl[w]) == so, w
        assert res.native_unconstrained_dcal[w] == oracle.mfe(frag, structure=False)[0]
        o = oracle.pf(frag)                          # Q6: the partition function runs before sc is added
        assert _close(res.ed[w], o["ed"]) and _close(res.ensemble_dG[w], o["dG"]), w
        assert db_from_pt(res.centroid_tbl[w]) == o["centroid"]


@pytest.mark.parametrize("W,n", [(300, 6), (600, 3)])
def test_c5_hard_constraints_long_windows(engine, oracle, W, n):
    """C5: 100-kb record, 'x' with p = 0.1 on line 3, 100 mono shuffles, W = 300 / 600"""
    L, r, w0 = 100000, 100, 50000
    seq = synth(L, 1005)
    hc = hard_constraints(L)
    res = engine.scan(seq, W, 1, r, seed=42, hc=hc, first_window=w0, n_windows=n, final_window=False,
                      keep_shuffles=True)
    shuf = ["".join(map(chr, res.shuffles[w, k])) for w in range(n) for k in range(r)]
    assert np.array_equal(_oracle_energies(oracle, shuf).reshape(n, r), res.shuffle_dcal)
    for w in range(n):
        frag, hcw = seq[w0 + w:w0 + w + W], hc[w0 + w:w0 + w + W]
        eo, so = oracle.mfe(frag, hc=hcw)
        assert res.mfe_dcal[w] == eo and db_from_pt(res.pair_tbl[w]) == so, (W, w)
        assert all(so[k] == "." for k in range(W) if hcw[k] == "x")
        assert res.native_unconstrained_dcal[w] == oracle.mfe(frag, structure=False)[0]
        o = oracle.pf(frag, hc=hcw)
        assert _close(res.ed[w], o["ed"]) and _close(res.ensemble_dG[w], o["dG"]), (W, w)
        assert db_from_pt(res.centroid_tbl[w]) == o["centroid"]


@pytest.mark.parametrize("W,n_seq", [(300, 5), (450, 4), (600, 4)])
def test_partition_function_long_windows(engine, oracle, W, n_seq):
    seqs = [synth(W, 7000 + W + k) for k in range(n_seq - 1)] + [("GGGGAAAACCCC" * W)[:W]]
    res = engine.pf_batch(seqs, want_bpp=True)
    for k, s in enumerate(seqs):
        o = oracle.pf(s, want_bpp=True)
        assert _close(res["dG"][k], o["dG"]) and _close(res["ed"][k], o["ed"]), (W, k)
        assert np.abs(res["bpp"][k] - o["bpp"]).max() < 1e-9
        assert db_from_pt(res["centroid"][k]) == o["centroid"]


def test_c4_chunk_boundary_w120_r100(engine, oracle):
    """A record with more than 65,536 windows at the C4 / C2 geometry (W=120, step 1, 100 shuffles): the windows either
    side of the chunk boundary equal a separate scan of that range (shuffles are keyed by the absolute window index)
    and the oracle."""
    W, r = 120, 100
    L = 65536 + 200 + W - 1
    seq = synth(L, 1004)
    full = engine.scan(seq, W, 1, r, seed=42)
    n = L - W + 1
    assert full.n == n + 1
    lo, cnt = 65536 - 8, 16
    part = engine.scan(seq, W, 1, r, seed=42, first_window=lo, n_windows=cnt, final_window=False, keep_shuffles=True)
    for key in ("mfe_dcal", "native_unconstrained_dcal", "shuffle_dcal", "pair_tbl", "centroid_tbl"):
        assert np.array_equal(getattr(full, key)[lo:lo + cnt], getattr(part, key)[:cnt]), key
    assert np.allclose(full.ed[lo:lo + cnt], part.ed[:cnt], rtol=0, atol=1e-12)
    for w in (6, 7, 8, 9):                             # 65534 .. 65537
        frag = seq[lo + w:lo + w + W]
        e, s = oracle.mfe(frag)
        assert e == part.mfe_dcal[w] and s == db_from_pt(part.pair_tbl[w])
        shuf = ["".join(map(chr, part.shuffles[w, k])) for k in range(r)]
        assert all(sorted(x) == sorted(frag) for x in shuf)
        assert np.array_equal(_oracle_energies(oracle, shuf), part.shuffle_dcal[w])
        assert _close(part.ed[w], oracle.pf(frag)["ed"])
    assert full.mfe_dcal[n] == full.mfe_dcal[n - 1]    # final-window slot: the stale fold compound (Q5)


# ------------------------------------------------------------------------------------------------ shuffle statistics
def _host_shuffles(seq, W, step, n, r, stype, seed):
    rng = random.Random(seed)
    out = np.zeros((n, r, W), dtype=np.uint8)
    for w in range(n):
        frag = seq[w * step:w * step + W]
        for k in range(r):
            sh = "".join(rng.sample(frag, W)) if stype == "mono" else dinucl_shuffle(frag, rng)
            out[w, k] = np.frombuffer(sh.encode(), dtype=np.uint8)
    return out


@pytest.mark.parametrize("stype", ["mono", "di"])
def test_zscore_distribution_device_vs_host_shuffles(engine, stype):
    """north_star: "the Philox-shuffled mode must produce statistically matching z-score distributions".  2,400 windows
    of a C2-like record; z from device shuffles (ScanFoldFunctions.py:800-802 / :255-277 replaced by shuffle.cu) against z
    from host shuffles fed in parity mode: two-sample Kolmogorov-Smirnov on the z-scores, and the mean of the paired
    differences must be zero within four standard errors."""
    from scipy import stats as sps
    from scanfold_b200 import stats
    W, step, r, n = 120, 10, 100, 2400
    seq = synth((n - 1) * step + W, 1002, (0.299, 0.184, 0.196, 0.321))
    dev = engine.scan(seq, W, step, r, shuffle_type=stype, seed=42, final_window=False, want_pf=False)
    host_sh = _host_shuffles(seq, W, step, n, r, stype, 4242)
    host = engine.scan(seq, W, step, r, shuffle_type=stype, parity_shuffles=host_sh, final_window=False, want_pf=False)
    assert np.array_equal(dev.native_unconstrained_dcal, host.native_unconstrained_dcal)
    z_dev, p_dev = stats.zscore_pvalue(dev.native_unconstrained_dcal, dev.shuffle_dcal)
    z_host, p_host = stats.zscore_pvalue(host.native_unconstrained_dcal, host.shuffle_dcal)
    ks = sps.ks_2samp(z_dev, z_host)
    assert ks.pvalue >= 1e-3, ks
    assert sps.ks_2samp(p_dev, p_host).pvalue >= 1e-3
    diff = z_dev - z_host
    assert abs(diff.mean()) < 4 * diff.std(ddof=1) / np.sqrt(n), (diff.mean(), diff.std())
    # pooled background energies, centred per window: the two shuffle generators draw from the same distribution
    c_dev = (dev.shuffle_dcal - dev.shuffle_dcal.mean(axis=1, keepdims=True)).ravel()
    c_host = (host.shuffle_dcal - host.shuffle_dcal.mean(axis=1, keepdims=True)).ravel()
    sub = np.random.default_rng(1).choice(len(c_dev), 20000, replace=False)   # KS assumes independent draws
    assert sps.ks_2samp(c_dev[sub], c_host[sub]).pvalue >= 1e-3


def test_single_window_energy_distribution_mono(engine):
    """20,000 device shuffles of one window against 20,000 random.sample shuffles: KS on the folded energies"""
    from scipy import stats as sps
    W, r = 120, 20000
    frag = synth(W, 77, (0.299, 0.184, 0.196, 0.321))
    dev = engine.scan(frag, W, 1, r, seed=7, final_window=False, want_pf=False)
    host_sh = _host_shuffles(frag, W, 1, 1, r, "mono", 99)
    host = engine.scan(frag, W, 1, r, parity_shuffles=host_sh, final_window=False, want_pf=False)
    assert sps.ks_2samp(dev.shuffle_dcal[0], host.shuffle_dcal[0]).pvalue >= 1e-3


def test_dinucleotide_shuffle_frequencies_match_reference(engine):
    """Outcome frequencies of the device Altschul-Erikson shuffle against the table drawn from the reference's own
    dinuclShuffle (tests/golden/make_shuffle_fixture.py): chi-square test of homogeneity per sequence."""
    from scipy import stats as sps
    table = json.load(open(os.path.join(GOLDEN, "dishuffle_counts.json")))
    for case in table:
        s, draws, ref = case["seq"], case["draws"], case["counts"]
        W = len(s)
        res = engine.scan(s, W, 1, draws, shuffle_type="di", seed=2026, final_window=False, want_pf=False,
                          keep_shuffles=True)
        got = {}
        for row in res.shuffles[0]:
            k = bytes(row).decode()
            got[k] = got.get(k, 0) + 1
        assert set(got) <= set(ref), (s, sorted(set(got) - set(ref))[:3])     # no outcome the reference cannot produce
        keys = sorted(ref)
        obs = np.array([[got.get(k, 0) for k in keys], [ref[k] for k in keys]], dtype=np.float64)
        chi2, p, dof, _ = sps.chi2_contingency(obs)
        assert p >= 1e-3, (s, chi2, dof, p)
