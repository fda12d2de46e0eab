"""GPU parity tests: CUDA engine (through the C-ABI) vs the CPU oracle on the same seeded inputs.

Bit-exact for integer work (MFE energies, structures, accumulators); 1e-6 relative for the fp64 partition
function (north_star tolerance).  The oracle itself is unpinned against ViennaRNA (see oracle/sf_oracle.h).
"""
import random

import numpy as np
import pytest

from util import db_from_pt, rand_seqs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,n_seq", [(5, 20), (12, 50), (40, 200), (120, 300), (200, 24), (300, 4), (600, 2)])
def test_mfe_energy_and_structure(engine, oracle, W, n_seq):
    seqs = rand_seqs(1000 + W, n_seq, W, gc_rich=True)
    e, pt = engine.fold_batch(seqs, structure=True)
    e2, _ = engine.fold_batch(seqs, structure=False)
    assert np.array_equal(e, e2)
    for k, s in enumerate(seqs):
        eo, so = oracle.mfe(s)
        assert e[k] == eo, (W, k, s)
        assert db_from_pt(pt[k]) == so, (W, k, s)


def test_mfe_adversarial(engine, oracle):
    seqs = ["G" * 60 + "C" * 60, "GC" * 60, "A" * 120, "GGGGAAAACCCC" * 10, "AU" * 60, "GU" * 60,
            "CUUCGG" * 20, "GCAACGC" * 17 + "A"]
    e, pt = engine.fold_batch(seqs, structure=True)
    for k, s in enumerate(seqs):
        eo, so = oracle.mfe(s)
        assert e[k] == eo and db_from_pt(pt[k]) == so, s


def _rand_hc(rng, n, with_brackets):
    hc = ["."] * n
    for k in range(n):
        r = rng.random()
        if r < 0.08:
            hc[k] = "x"
        elif r < 0.11:
            hc[k] = "<"
        elif r < 0.14:
            hc[k] = ">"
        elif r < 0.16:
            hc[k] = "|"
    if with_brackets:
        for _ in range(3):
            i = rng.randrange(0, n - 10)
            j = rng.randrange(i + 5, n)
            inside = hc[i:j + 1]
            if "(" in inside or ")" in inside:
                continue
            # keep brackets nested: only place if no bracket lies between
            hc[i], hc[j] = "(", ")"
    return "".join(hc)


@pytest.mark.parametrize("brackets", [False, True])
def test_mfe_hard_constraints(engine, oracle, brackets):
    rng = random.Random(7 + brackets)
    W = 80
    seqs = rand_seqs(77, 60, W)
    hcs = [_rand_hc(rng, W, brackets) for _ in seqs]
    if brackets:
        hcs[0] = "((((" + "." * (W - 8) + "))))"
        hcs[1] = ")" + "." * (W - 2) + "("      # unbalanced: ignored
        hcs[2] = "((" + "." * (W - 3) + ")"     # one unmatched '('
    e, pt = engine.fold_batch(seqs, hc=hcs, structure=True)
    for k, s in enumerate(seqs):
        eo, so = oracle.mfe(s, hc=hcs[k])
        assert e[k] == eo, (k, s, hcs[k])
        assert db_from_pt(pt[k]) == so, (k, s, hcs[k])


def test_mfe_soft_constraints_and_span(engine, oracle):
    rng = np.random.default_rng(5)
    W = 100
    seqs = rand_seqs(55, 40, W)
    sc = rng.integers(-150, 150, size=(len(seqs), W + 1)).astype(np.int32)
    sc[:, 0] = 0
    e, pt = engine.fold_batch(seqs, sc=sc, structure=True)
    for k, s in enumerate(seqs):
        eo, so = oracle.mfe(s, sc_stack=sc[k])
        assert e[k] == eo and db_from_pt(pt[k]) == so, (k, s)
    e, pt = engine.fold_batch(seqs, structure=True, max_span=30)
    for k, s in enumerate(seqs):
        eo, so = oracle.mfe(s, max_span=30)
        assert e[k] == eo and db_from_pt(pt[k]) == so, (k, s)


def test_deigan_conversion(engine, oracle):
    rng = np.random.default_rng(9)
    r = np.concatenate([[-999.0], rng.exponential(0.4, size=300)])
    r[rng.random(301) < 0.05] = -999.0
    assert np.array_equal(engine.deigan(r, 0.8, -0.2), oracle.deigan(r, 0.8, -0.2))


@pytest.mark.parametrize("W,n_seq", [(30, 20), (120, 24), (200, 4)])
def test_partition_function(engine, oracle, W, n_seq):
    seqs = rand_seqs(2000 + W, n_seq, W, gc_rich=True)
    res = engine.pf_batch(seqs, want_bpp=True)
    for k, s in enumerate(seqs):
        o = oracle.pf(s, want_bpp=True)
        assert abs(res["dG"][k] - o["dG"]) <= 1e-6 * max(1.0, abs(o["dG"])), (k, s)
        assert abs(res["ed"][k] - o["ed"]) <= 1e-6 * max(1.0, abs(o["ed"])), (k, s)
        assert np.abs(res["bpp"][k] - o["bpp"]).max() < 1e-9
        assert db_from_pt(res["centroid"][k]) == o["centroid"]


def test_partition_function_constraints(engine, oracle):
    rng = random.Random(3)
    W = 90
    seqs = rand_seqs(31, 12, W)
    hcs = [_rand_hc(rng, W, k % 2 == 0) for k in range(len(seqs))]
    sc = np.random.default_rng(1).integers(-100, 100, size=(len(seqs), W + 1)).astype(np.int32)
    res = engine.pf_batch(seqs, hc=hcs, sc=sc, max_span=50)
    for k, s in enumerate(seqs):
        o = oracle.pf(s, hc=hcs[k], sc_stack=sc[k], max_span=50)
        assert abs(res["dG"][k] - o["dG"]) <= 1e-6 * max(1.0, abs(o["dG"]))
        assert abs(res["ed"][k] - o["ed"]) <= 1e-6 * max(1.0, abs(o["ed"]))
        assert db_from_pt(res["centroid"][k]) == o["centroid"]


@pytest.mark.parametrize("W", [40, 90, 120, 200])
def test_flag_only_hard_constraints_fast_kernels(engine, oracle, W):
    """Constraint lines without brackets ('x' '<' '>' '|' only) run on the fast kernels (mfe3 with traceback, pf2 up to
    120 nt): energies / structures bit-exact and PF within 1e-6 of the oracle, and equal to the int32 / first PF kernel."""
    rng = random.Random(100 + W)
    seqs = rand_seqs(900 + W, 48, W, gc_rich=True)
    hcs = [_rand_hc(rng, W, False) for _ in seqs]
    hcs[0] = "x" * W
    hcs[1] = "." * W
    e, pt = engine.fold_batch(seqs, hc=hcs, structure=True)
    res = engine.pf_batch(seqs, hc=hcs)
    try:
        engine.set_engines(mfe=1, pf=1)
        e1, pt1 = engine.fold_batch(seqs, hc=hcs, structure=True)
        res1 = engine.pf_batch(seqs, hc=hcs)
    finally:
        engine.set_engines(mfe=3, pf=2)
    assert np.array_equal(e, e1) and np.array_equal(pt, pt1)
    assert np.allclose(res["ed"], res1["ed"], rtol=1e-9, atol=1e-9) and np.allclose(res["dG"], res1["dG"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(res["centroid"], res1["centroid"])
    for k in range(0, len(seqs), 3):
        eo, so = oracle.mfe(seqs[k], hc=hcs[k])
        assert e[k] == eo and db_from_pt(pt[k]) == so, (W, k, seqs[k], hcs[k])
        o = oracle.pf(seqs[k], hc=hcs[k])
        assert abs(res["dG"][k] - o["dG"]) <= 1e-6 * max(1.0, abs(o["dG"]))
        assert abs(res["ed"][k] - o["ed"]) <= 1e-6 * max(1.0, abs(o["ed"]))
        assert db_from_pt(res["centroid"][k]) == o["centroid"]


def test_random_parameter_file(engine, oracle, tmp_path):
    """A randomised table set catches any index-order disagreement between the two table loaders / kernels."""
    import os
    from scanfold_b200 import engine as E
    src = open(os.path.join(os.path.dirname(E.LIB_PATH), "params", "rna_turner2004_besteffort.par")).read().split("\n")
    rng = random.Random(11)
    out, in_block, name = [], False, ""
    for line in src:
        if line.startswith("# "):
            name = line[2:].strip()
            out.append(line)
            continue
        toks = line.split()
        numeric = toks and all(t.lstrip("-").isdigit() or t == "INF" for t in toks)
        if numeric and name not in ("hairpin", "bulge", "interior", "ML_params", "NINIO", "Misc") and not name.endswith("_enthalpies"):
            out.append(" ".join(str(rng.randrange(-300, 300)) for _ in toks))
        else:
            out.append(line)
    par = tmp_path / "random.par"
    par.write_text("\n".join(out))
    try:
        E.init(0, str(par))
        oracle.load_params(str(par))
        seqs = rand_seqs(99, 80, 70)
        e, pt = E.fold_batch(seqs, structure=True)
        for k, s in enumerate(seqs):
            eo, so = oracle.mfe(s)
            assert e[k] == eo and db_from_pt(pt[k]) == so, (k, s)
        res = E.pf_batch(seqs[:10])
        for k in range(10):
            o = oracle.pf(seqs[k])
            assert abs(res["ed"][k] - o["ed"]) <= 1e-6 * max(1.0, abs(o["ed"]))
    finally:
        E.init(0, None)
        oracle.load_params(oracle.DEFAULT_PAR)


def _dinuc_counts(s):
    c = {}
    for a, b in zip(s[:-1], s[1:]):
        c[a + b] = c.get(a + b, 0) + 1
    return c


@pytest.mark.parametrize("stype", ["mono", "di"])
def test_device_shuffles_invariants(engine, stype):
    seq = rand_seqs(4242, 1, 400)[0]
    W, r = 60, 25
    plan = engine.ScanPlan(seq, W, 7, r, shuffle_type=stype, seed=123, want_pf=False, keep_shuffles=True)
    res = plan.run().fetch()
    plan.close()
    n_reg = res.n - 1
    distinct = set()
    for w in range(res.n):
        start = w * 7 if w < n_reg else len(seq) - W
        frag = seq[start:start + W]
        for k in range(r):
            sh = bytes(res.shuffles[w, k]).decode()
            distinct.add(sh)
            assert sorted(sh) == sorted(frag)
            if stype == "di":
                assert sh[0] == frag[0] and sh[-1] == frag[-1]
                assert _dinuc_counts(sh) == _dinuc_counts(frag)
    assert len(distinct) > 0.95 * res.n * r
    # same seed -> same shuffles regardless of sharding (Philox counter = absolute window index)
    plan2 = engine.ScanPlan(seq, W, 7, r, shuffle_type=stype, seed=123, want_pf=False, keep_shuffles=True,
                            first_window=10, n_windows=5, final_window=False)
    res2 = plan2.run().fetch()
    plan2.close()
    assert np.array_equal(res2.shuffles, res.shuffles[10:15])
    assert np.array_equal(res2.shuffle_dcal, res.shuffle_dcal[10:15])


def test_mono_shuffle_uniformity(engine):
    """position of the first nucleotide after shuffling is uniform (chi-square, generous bound)"""
    W, r = 24, 100
    seq = "A" + "C" * (W - 1)
    counts = np.zeros(W)
    for seed in range(40):
        p = engine.ScanPlan(seq[:W], W, 1, r, seed=seed, want_pf=False, keep_shuffles=True, final_window=False)
        sh = p.run().fetch().shuffles[0]
        p.close()
        for k in range(r):
            counts[bytes(sh[k]).index(b"A")] += 1
    exp = counts.sum() / W
    chi2 = ((counts - exp) ** 2 / exp).sum()
    assert chi2 < 60, chi2   # 23 dof: P(chi2 > 60) ~ 3e-5


def test_scan_parity_mode(engine, oracle):
    """sfb_scan with host-provided shuffles vs the oracle, all outputs (the scan loop ScanFold.py:429-757)."""
    rng = random.Random(17)
    seq = rand_seqs(808, 1, 230)[0]
    W, step, r = 50, 3, 12
    nwin = (len(seq) - W) // step + 1
    shuf = np.zeros((nwin + 1, r, W), dtype=np.uint8)
    for w in range(nwin + 1):
        start = w * step if w < nwin else len(seq) - W
        frag = seq[start:start + W]
        for k in range(r):
            shuf[w, k] = np.frombuffer("".join(rng.sample(frag, W)).encode(), dtype=np.uint8)
    hc = "".join(rng.choice("....x") for _ in seq)
    res = engine.scan(seq, W, step, r, parity_shuffles=shuf, hc=hc)
    assert res.n == nwin + 1
    for w in range(nwin + 1):
        start = w * step if w < nwin else len(seq) - W
        frag = seq[start:start + W]
        src = min(w, nwin - 1)          # final slot: stale fold compound of the last regular window (Q5)
        sfrag = seq[src * step:src * step + W]
        shc = hc[src * step:src * step + W]
        eo, so = oracle.mfe(sfrag, hc=shc)
        assert res.mfe_dcal[w] == eo
        assert db_from_pt(res.pair_tbl[w]) == so
        assert res.native_unconstrained_dcal[w] == oracle.mfe(frag, structure=False)[0]
        for k in range(r):
            assert res.shuffle_dcal[w, k] == oracle.mfe(bytes(shuf[w, k]).decode(), structure=False)[0]
        o = oracle.pf(sfrag, hc=shc)
        assert abs(res.ed[w] - o["ed"]) <= 1e-6 * max(1, abs(o["ed"]))
        assert db_from_pt(res.centroid_tbl[w]) == o["centroid"]


def test_scan_shape_mode(engine, oracle):
    rng = np.random.default_rng(21)
    seq = rand_seqs(909, 1, 160)[0]
    L, W, step, r = len(seq), 40, 1, 3
    react = np.concatenate([[-999.0], np.clip(rng.exponential(0.4, L), 0, 4)])
    react[rng.random(L + 1) < 0.05] = -999.0
    res = engine.scan(seq, W, step, r, react=react, shape_m=0.8, shape_b=-0.2)
    es = oracle.deigan(react, 0.8, -0.2)
    nwin = L - W + 1
    for w in range(nwin):
        frag = seq[w:w + W]
        scw = np.zeros(W + 1, dtype=np.int32)
        for p in range(1, W):                       # Q7: +1 shift, position W reads past the slice
            g = w + 1 + p
            scw[p] = es[g] if g <= L else 0
        eo, so = oracle.mfe(frag, sc_stack=scw)
        assert res.mfe_dcal[w] == eo and db_from_pt(res.pair_tbl[w]) == so
        o = oracle.pf(frag)                          # Q6: PF before sc is added
        assert abs(res.ed[w] - o["ed"]) <= 1e-6 * max(1, abs(o["ed"]))
        if w == nwin - 1:                            # Q5: final slot re-runs pf() on the stale fc WITH sc
            o2 = oracle.pf(frag, sc_stack=scw)
            assert abs(res.ed[nwin] - o2["ed"]) <= 1e-6 * max(1, abs(o2["ed"]))
            assert res.mfe_dcal[nwin] == eo


@pytest.mark.parametrize("W", [16, 17, 33, 40, 64, 65, 97, 119, 120, 121, 128, 150, 199, 200, 201, 250, 299, 300, 301])
def test_fold_kernels_agree(engine, oracle, W):
    """The three MFE kernel generations -- int32 CTA kernel (mfe.cu), int16 warp teams (mfe2.cu), int16 CTA kernel
    with stencil / range-minimum interior loops and on-device traceback (mfe3.cu) -- give the same energies and
    structures, bit for bit, including sequences that overflow int16 and get redone, and equal the oracle."""
    seqs = rand_seqs(7000 + W, 600, W, gc_rich=True)
    seqs += ["G" * (W // 2) + "C" * (W - W // 2), ("GC" * W)[:W], "A" * W, ("GGGGAAAACCCC" * W)[:W], ("AU" * W)[:W],
             ("GGGGGCCCCC" * W)[:W], ("CUUCGG" * W)[:W], "N" * W, ("ACGUN" * W)[:W]]
    try:
        engine.set_engines(mfe=1)
        e_ref, pt_ref = engine.fold_batch(seqs, structure=True)
        engine.set_engines(mfe=2)
        e2, _ = engine.fold_batch(seqs, structure=False)
        engine.set_engines(mfe=3)
        e3, _ = engine.fold_batch(seqs, structure=False)
        e3s, pt3 = engine.fold_batch(seqs, structure=True)
    finally:
        engine.set_engines(mfe=3)
    for name, e in (("mfe2", e2), ("mfe3", e3), ("mfe3+traceback", e3s)):
        bad = np.nonzero(e != e_ref)[0]
        assert len(bad) == 0, (name, W, [(int(k), seqs[k], int(e[k]), int(e_ref[k])) for k in bad[:5]])
    bad = np.nonzero((pt3 != pt_ref).any(axis=1))[0]
    assert len(bad) == 0, (W, [(int(k), seqs[k]) for k in bad[:5]])
    for k in list(range(0, 40)) + list(range(len(seqs) - 9, len(seqs))):
        eo, so = oracle.mfe(seqs[k])
        assert e3[k] == eo and db_from_pt(pt3[k]) == so, (W, k, seqs[k])


@pytest.mark.parametrize("W", [10, 31, 64, 97, 120])
def test_partition_function_kernels_agree(engine, oracle, W):
    """Shared-memory PF kernel (pf2.cu) against the global-memory one (pf.cu) and the oracle."""
    seqs = rand_seqs(4100 + W, 40, W, gc_rich=True) + ["A" * W, ("GC" * W)[:W], ("GGGGAAAACCCC" * W)[:W]]
    try:
        engine.set_engines(pf=1)
        r1 = engine.pf_batch(seqs, want_bpp=True)
        engine.set_engines(pf=2)
        r2 = engine.pf_batch(seqs, want_bpp=True)
    finally:
        engine.set_engines(pf=2)
    for key in ("dG", "ed"):
        assert np.allclose(r1[key], r2[key], rtol=1e-9, atol=1e-9), key
    assert np.abs(r1["bpp"] - r2["bpp"]).max() < 1e-10
    assert np.array_equal(r1["centroid"], r2["centroid"])
    for k in range(0, len(seqs), 7):
        o = oracle.pf(seqs[k], want_bpp=True)
        assert abs(r2["dG"][k] - o["dG"]) <= 1e-6 * max(1.0, abs(o["dG"]))
        assert abs(r2["ed"][k] - o["ed"]) <= 1e-6 * max(1.0, abs(o["ed"]))
        assert np.abs(r2["bpp"][k] - o["bpp"]).max() < 1e-9


def test_scan_across_chunk_boundary(engine, oracle):
    """A record with more than 65,536 windows is scanned in chunks: the windows on both sides of the boundary (and the
    final-window slot) must equal a separate scan of just that range -- device shuffles are keyed by the absolute window
    index -- and the oracle."""
    rng = np.random.default_rng(21)
    W, r = 40, 4
    L = 65536 + 300 + W - 1
    seq = "".join("ACGU"[k] for k in rng.integers(0, 4, L))
    full = engine.scan(seq, W, 1, r, seed=3, keep_shuffles=True)
    n = L - W + 1
    assert full.n == n + 1
    lo, cnt = 65536 - 20, 40
    part = engine.scan(seq, W, 1, r, seed=3, first_window=lo, n_windows=cnt, final_window=False, keep_shuffles=True)
    for key in ("mfe_dcal", "native_unconstrained_dcal", "shuffle_dcal", "pair_tbl", "centroid_tbl", "shuffles"):
        assert np.array_equal(getattr(full, key)[lo:lo + cnt], getattr(part, key)[:cnt]), key
    assert np.allclose(full.ed[lo:lo + cnt], part.ed[:cnt], rtol=0, atol=1e-12)
    for w in (65535, 65536, n - 1):
        frag = seq[w:w + W]
        e, s = oracle.mfe(frag)
        assert e == full.mfe_dcal[w] and s == db_from_pt(full.pair_tbl[w])
        for k in range(r):
            sh = bytes(full.shuffles[w, k]).decode()
            assert sorted(sh) == sorted(frag) and oracle.mfe(sh, structure=False)[0] == full.shuffle_dcal[w, k]
    # final-window slot: the stale fold compound of the last regular window, fresh shuffles of seq[L-W:L] (Q5)
    assert full.mfe_dcal[n] == full.mfe_dcal[n - 1] and np.array_equal(full.pair_tbl[n], full.pair_tbl[n - 1])


@pytest.mark.parametrize("W", [301, 333, 450, 600, 1024])
def test_blocked_kernel_long_windows(engine, oracle, W):
    """Windows above 300 nt run on the blocked int32 kernel (mfe4.cu): energies and structures equal the int32 CTA kernel
    (mfe.cu, selected with set_engines(mfe=1)) and the oracle -- unconstrained, with per-nucleotide flags, with enforced
    pairs, with stacking pseudo-energies and with a span limit."""
    rng = random.Random(W)
    n_seq = 10 if W <= 600 else 4
    seqs = rand_seqs(8100 + W, n_seq, W, gc_rich=True) + [("GGGGAAAACCCC" * W)[:W], ("GC" * W)[:W], "A" * W, ("ACGUN" * W)[:W]]
    hc_flags = [_rand_hc(rng, W, False) for _ in seqs]
    hc_br = [_rand_hc(rng, W, True) for _ in seqs]
    hc_br[0] = "((((" + "." * (W - 8) + "))))"
    hc_br[1] = ")" + "." * (W - 2) + "("
    sc = np.random.default_rng(W).integers(-150, 150, size=(len(seqs), W + 1)).astype(np.int32)
    sc[:, 0] = 0
    cases = {"plain": {}, "flags": {"hc": hc_flags}, "brackets": {"hc": hc_br}, "sc": {"sc": sc}, "span": {"max_span": 150}}
    got, ref = {}, {}
    for name, kw in cases.items():
        got[name] = engine.fold_batch(seqs, structure=True, **kw)
    e_only, _ = engine.fold_batch(seqs, structure=False)
    assert np.array_equal(e_only, got["plain"][0])
    try:
        engine.set_engines(mfe=1)
        for name, kw in cases.items():
            ref[name] = engine.fold_batch(seqs, structure=True, **kw)
    finally:
        engine.set_engines(mfe=3)
    for name in cases:
        bad = np.nonzero(got[name][0] != ref[name][0])[0]
        assert len(bad) == 0, (name, W, [(int(k), int(got[name][0][k]), int(ref[name][0][k])) for k in bad[:5]])
        assert np.array_equal(got[name][1], ref[name][1]), (name, W)
    for k in (0, len(seqs) - 4):
        for name, okw in (("plain", {}), ("flags", {"hc": hc_flags[k]}), ("brackets", {"hc": hc_br[k]}),
                          ("sc", {"sc_stack": sc[k]}), ("span", {"max_span": 150})):
            if W > 600 and name not in ("plain", "brackets"):
                continue
            eo, so = oracle.mfe(seqs[k], **okw)
            assert got[name][0][k] == eo and db_from_pt(got[name][1][k]) == so, (name, W, k)


def _nested_constraint(rng, n, n_pairs):
    """a constraint line with nested / side-by-side enforced pairs '(' ')' plus a few 'x' flags"""
    hc = ["."] * n
    free = [(0, n - 1)]
    for _ in range(n_pairs):
        if not free:
            break
        lo, hi = free.pop(rng.randrange(len(free)))
        if hi - lo < 12:
            continue
        i = rng.randrange(lo, lo + (hi - lo) // 3)
        j = rng.randrange(hi - (hi - lo) // 3, hi + 1)
        if j - i < 5:
            continue
        hc[i], hc[j] = "(", ")"
        free += [(lo, i - 1), (i + 1, j - 1), (j + 1, hi)]
    for k in range(n):
        if hc[k] == "." and rng.random() < 0.03:
            hc[k] = "x"
    return "".join(hc)


@pytest.mark.parametrize("n", [70, 700, 2000])
def test_whole_sequence_fold(engine, oracle, n):
    """sfb_fold_long (the full-length folds of --global_refold, ScanFold.py:1518-1539): unconstrained and with a dbn-like
    constraint line of enforced pairs, against the oracle -- energy and structure, 32-bit pair table."""
    rng = random.Random(n)
    seq = rand_seqs(31337 + n, 1, n)[0]
    hc = _nested_constraint(rng, n, 60)
    for h in (None, hc, hc[:n - 7]):                 # a shorter line constrains the leading positions only
        e, pt = engine.fold_long(seq, hc=h)
        full = None if h is None else h + "." * (n - len(h))
        eo, so = oracle.mfe(seq, hc=full)
        assert e == eo, (n, h is not None)
        assert db_from_pt(pt) == so
        if h is not None:
            for k, ch in enumerate(full):
                if ch == "x":
                    assert pt[k] == 0


@pytest.mark.parametrize("T", [25.0, 50.0, 10.0])
def test_temperature_rescaling(engine, oracle, T):
    """md.temperature != 37 (ScanFold.py:212-213, -t): every table rescaled from the 37 C values and the enthalpy blocks of
    the parameter file, on all MFE kernels (short, 120-nt, 300-nt, blocked) and both PF kernels, against the oracle."""
    try:
        for W, n_seq in ((40, 40), (120, 40), (200, 6), (320, 4)):
            seqs = rand_seqs(int(T) * 1000 + W, n_seq, W, gc_rich=True)
            e, pt = engine.fold_batch(seqs, structure=True, temperature=T)
            e37, _ = engine.fold_batch(seqs, structure=False)
            assert not np.array_equal(e, e37)
            for k in range(0, n_seq, 3):
                eo, so = oracle.mfe(seqs[k], temperature=T)
                assert e[k] == eo and db_from_pt(pt[k]) == so, (T, W, k)
            if W <= 200:
                res = engine.pf_batch(seqs[:6], temperature=T)
                for k in range(0, 6, 2):
                    o = oracle.pf(seqs[k], temperature=T)
                    assert abs(res["dG"][k] - o["dG"]) <= 1e-6 * max(1.0, abs(o["dG"])), (T, W, k)
                    assert abs(res["ed"][k] - o["ed"]) <= 1e-6 * max(1.0, abs(o["ed"])), (T, W, k)
                    assert db_from_pt(res["centroid"][k]) == o["centroid"]
        # back at 37 the tables are the file's own values again
        s = rand_seqs(5, 3, 90)
        e, _ = engine.fold_batch(s, structure=False)
        assert [int(x) for x in e] == [oracle.mfe(x, structure=False)[0] for x in s]
    finally:
        oracle.set_temperature(37.0)


def test_scan_record_in_pipelined_parts_equals_one_call(engine, monkeypatch):
    """scan.scan_record splits long shards into parts (host statistics of part k overlap the folds of part k+1): same
    table as one engine call, device shuffles (Philox by absolute window) and parity shuffles alike, final-window set
    included, also for a window range that starts inside the record"""
    from scanfold_b200 import scan
    rng = np.random.default_rng(17)
    seq = "".join("ACGU"[k] for k in rng.integers(0, 4, 420))
    W, step, r = 40, 3, 6
    total = scan.n_windows_of(len(seq), W, step)
    par = rng.integers(0, 4, (total + 1, r, W)).astype(np.uint8)
    par = np.frombuffer(b"ACGU", dtype=np.uint8)[par]
    fields = ("start1", "end1", "mfe_dcal", "mfe", "z", "p", "ed", "native_unconstrained_dcal", "shuffle_dcal", "pair_tbl", "centroid_tbl")
    for kw in (dict(), dict(parity_shuffles=par), dict(first_window=7, n_windows=total - 7, shuffle_type="di"),
               dict(first_window=5, n_windows=60)):
        monkeypatch.setattr(scan, "PIPELINE_WINDOWS", 1 << 30)
        one = scan.scan_record(seq, W, step, r, seed=9, **kw)
        monkeypatch.setattr(scan, "PIPELINE_WINDOWS", 37)
        parts = scan.scan_record(seq, W, step, r, seed=9, **kw)
        assert len(parts) == len(one)
        for f in fields:
            assert np.array_equal(getattr(one, f), getattr(parts, f)), (f, kw.keys())
        assert one.final == parts.final
