"""The CPU oracle checked against itself by independent routes (it has no ViennaRNA vector to be pinned to --
"parity unpinned", oracle/sf_oracle.h): exhaustive enumeration on short sequences, an independent loop-energy
evaluator, partition-function identities, and the work counters against closed forms."""
import itertools
import math
import random

import numpy as np
import pytest

from util import rand_seqs

PAIRS = {("C", "G"), ("G", "C"), ("G", "U"), ("U", "G"), ("A", "U"), ("U", "A")}


def all_structures(seq):
    """every secondary structure (canonical pairs, hairpins >= 3) of a short sequence, as dot-bracket"""
    n = len(seq)

    def rec(i, j):
        if j - i < 4:
            return [[]]
        out = [s for s in rec(i + 1, j)]
        for k in range(i + 4, j + 1):
            if (seq[i], seq[k]) in PAIRS:
                for a in rec(i + 1, k - 1):
                    for b in rec(k + 1, j):
                        out.append([(i, k)] + a + b)
        return out

    res = []
    for pairs in rec(0, n - 1):
        s = ["."] * n
        for a, b in pairs:
            s[a], s[b] = "(", ")"
        res.append("".join(s))
    return res


@pytest.mark.parametrize("seed", range(6))
def test_mfe_is_the_minimum_over_all_structures(oracle, seed):
    rng = random.Random(seed)
    seq = "".join(rng.choice("ACGU" if seed % 2 else "GGCCAU") for _ in range(14 + seed % 3))
    structs = all_structures(seq)
    energies = [oracle.eval_structure(seq, s) for s in structs]
    e, s = oracle.mfe(seq)
    assert e == min(energies)
    assert oracle.eval_structure(seq, s) == e


def test_traceback_energy_equals_dp_value(oracle):
    for s in rand_seqs(5, 60, 90, gc_rich=True):
        e, db = oracle.mfe(s)
        assert oracle.eval_structure(s, db) == e
        assert db.count("(") == db.count(")")


def test_partition_function_by_enumeration(oracle):
    """Z from the DP equals the Boltzmann sum over all structures (same smoothed dangle model)"""
    kT = (37 + 273.15) * 1.98717 / 1000.0
    for seed in range(4):
        rng = random.Random(100 + seed)
        seq = "".join(rng.choice("GGCCAU") for _ in range(15))
        z = sum(oracle.eval_weight(seq, s) for s in all_structures(seq))
        o = oracle.pf(seq, want_bpp=True)
        assert abs(o["dG"] - (-kT * math.log(z))) < 1e-9 * max(1.0, abs(o["dG"]))
        # base-pair probabilities from enumeration
        n = len(seq)
        bpp = np.zeros((n, n))
        for s in all_structures(seq):
            w = oracle.eval_weight(seq, s) / z
            stk = []
            for k, ch in enumerate(s):
                if ch == "(":
                    stk.append(k)
                elif ch == ")":
                    bpp[stk.pop(), k] += w
        assert np.abs(bpp - o["bpp"]).max() < 1e-9
        ed = 2 * (bpp * (1 - bpp)).sum()
        assert abs(ed - o["ed"]) < 1e-9


def test_pf_identities(oracle):
    kT = (37 + 273.15) * 1.98717 / 1000.0
    for s in rand_seqs(9, 10, 80):
        e, _ = oracle.mfe(s)
        o = oracle.pf(s, want_bpp=True)
        assert o["dG"] <= e / 100.0 + 1e-9          # Z >= exp(-MFE/kT)
        p = o["bpp"] + o["bpp"].T
        assert p.sum(axis=1).max() <= 1 + 1e-9
        assert o["ed"] >= 0


def test_constraints_only_raise_the_energy(oracle):
    rng = random.Random(3)
    for s in rand_seqs(12, 20, 70):
        e0, _ = oracle.mfe(s)
        hc = "".join(rng.choice("....x<>") for _ in s)
        e1, db = oracle.mfe(s, hc=hc)
        assert e1 >= e0
        for k, ch in enumerate(hc):
            if ch == "x":
                assert db[k] == "."
            if ch == "<":
                assert db[k] != ")"
            if ch == ">":
                assert db[k] != "("
        e2, db2 = oracle.mfe(s, max_span=25)
        assert e2 >= e0
        stk = []
        for k, ch in enumerate(db2):
            if ch == "(":
                stk.append(k)
            elif ch == ")":
                assert k - stk.pop() + 1 <= 25


def test_work_counters_match_closed_forms(oracle):
    from scanfold_b200 import workcount as wc
    assert [wc.dense_relaxations(w) for w in (40, 120, 200)] == [72508, 2475988, 8783468]
    assert [wc.cells(w) for w in (40, 120, 200, 300, 600)] == [666, 6786, 19306, 43956, 177906]
    for s in rand_seqs(4, 3, 100):
        oracle.mfe(s)
        dense, useful = oracle.counters()
        assert dense == wc.dense_relaxations(len(s))
        assert useful == wc.useful_relaxations(s)


@pytest.mark.parametrize("W", [4, 5, 9, 17, 40, 77, 120, 200])
def test_fast_batch_path_equals_simple(oracle, W):
    """The tuned batch path bench.py's CPU arm times (sfo_fold_batch_fast: reusable buffers, vectorised sweeps over
    mismatch-carrying copies of C) returns exactly the energies of the simple checker path, sequence by sequence."""
    from util import rand_seqs
    n = 200 if W <= 120 else 24
    seqs = rand_seqs(31 * W, n, W, gc_rich=True) + ["A" * W, ("GC" * W)[:W], ("GGGGAAAACCCC" * W)[:W], "N" * W,
                                                    ("ACGUN" * W)[:W], ("CUUCGG" * W)[:W], ("gggaaaccc" * W)[:W]]
    a = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).reshape(len(seqs), W)
    simple = oracle.fold_batch(a, n_threads=2)
    fast = oracle.fold_batch(a, n_threads=3, fast=True)
    assert np.array_equal(simple, fast)
    for k in (0, 1, len(seqs) - 1):
        assert simple[k] == oracle.mfe(seqs[k], structure=False)[0]
