"""Level-1 drop-in on the GPU: `import RNA` resolving to scanfold_b200/compat/RNA.py with the CUDA backend, driven with
the exact call sequence of the reference's window loop (ScanFold.py:494-544, ScanFoldFunctions.py:774-789), must return
what the golden runs of the unmodified reference recorded (tests/golden/*/trace.npz; those runs used the same shim over
the CPU oracle).  /root/reference itself is not on the GPU box, so its loop is restated here call for call."""
import importlib
import json
import os
import sys

import numpy as np
import pytest

from test_host_pipeline import GOLDEN, load_case
from util import db_from_pt

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def RNA(engine, monkeypatch):
    """the module a reference install would import: scanfold_b200/compat on sys.path, CUDA backend"""
    monkeypatch.syspath_prepend(os.path.join(ROOT, "scanfold_b200", "compat"))
    sys.modules.pop("RNA", None)
    mod = importlib.import_module("RNA")
    from scanfold_b200 import rna_shim
    rna_shim.set_backend(rna_shim.EngineBackend())
    assert mod.__version__.startswith("scanfold_b200")
    yield mod
    sys.modules.pop("RNA", None)


def _read_react(path):
    from scanfold_b200.cli import read_reactivities
    return read_reactivities(path)


@pytest.mark.parametrize("name", ["mono_w40", "hc_w40", "shape_w40", "dna_name_w30"])
def test_reference_window_loop_through_the_shim(RNA, name):
    case = load_case(name)
    t = case["trace"]
    seq, W, step, n = case["seq"], case["W"], case["step"], case["n_windows"]
    args = case["args"]
    constraints = open(os.path.join(case["dir"], "constraints.dbn")).readlines()[2] if "--constraints" in args else None
    react = _read_react(os.path.join(case["dir"], "react.shape")) if "--react" in args else None
    md = RNA.md()
    md.temperature = 37
    for w in range(0, n, 3):
        i = w * step
        frag = seq[i:i + W]
        start, end = i + 1, i + W
        fc = RNA.fold_compound(str(frag), md)
        if constraints is None and react is None:                  # ScanFold.py:496-504
            structure, mfe = fc.mfe()
            fc.pf()
            centroid, _ = fc.centroid()
            ed = round(fc.mean_bp_distance(), 2)
        elif constraints is not None:                               # :508-519
            fc.hc_add_from_db("".join(list(constraints)[start - 1:end]))
            structure, mfe = fc.mfe()
            fc.pf()
            centroid, _ = fc.centroid()
            ed = round(fc.mean_bp_distance(), 2)
        else:                                                       # :522-541 (partition function before the pseudo-energies)
            fc.pf()
            centroid, _ = fc.centroid()
            ed = round(fc.mean_bp_distance(), 2)
            fc.sc_add_SHAPE_deigan(react[start:end + 1], 0.8, -0.2)
            structure, mfe = fc.mfe()
        assert structure == db_from_pt(t["pair_tbl"][w]), (name, w)
        assert mfe == float(np.float32(t["mfe_dcal"][w] / 100.0)), (name, w)
        assert centroid == db_from_pt(t["centroid_tbl"][w]), (name, w)
        assert ed == round(float(t["ed"][w]), 2), (name, w)
        # rna_folder (ScanFoldFunctions.py:774-789): a fresh md with the temperature only, energy of native + shuffles
        for k in (0, 1):
            md2 = RNA.md()
            md2.temperature = 37
            shuffled = bytes(t["shuffles"][w, k]).decode()
            _, e = RNA.fold_compound(shuffled, md2).mfe()
            assert e == float(np.float32(t["shuffle_dcal"][w, k] / 100.0)), (name, w, k)
        md3 = RNA.md()
        md3.temperature = 37
        assert RNA.fold_compound(str(frag), md3).mfe()[1] == float(np.float32(t["native_unconstrained_dcal"][w] / 100.0))


def test_motif_calls_through_the_shim(RNA):
    """RNA.fold / RNA.pf_fold and the constrained fold compound of the motif step (ScanFold.py:1733-1741)"""
    seq = "GGGGCGCUUCGGCGCCCCAUAAUUAAUA"
    s, e = RNA.fold(seq)
    assert len(s) == len(seq) and s.count("(") == s.count(")") and e < 0
    cen, eg = RNA.pf_fold(seq)
    assert len(cen) == len(seq) and eg <= e + 1e-6
    fc = RNA.fold_compound(seq)
    fc.hc_add_from_db("((((((((....))))))))........")
    s2, e2 = fc.mfe()
    assert len(s2) == len(seq) and e2 >= e - 1e-6          # weak enforcement: conflicting pairs removed, nothing forced
    pairs, stack = set(), []
    for k, ch in enumerate(s2):
        if ch == "(":
            stack.append(k)
        elif ch == ")":
            pairs.add((stack.pop(), k))
    for i, j in pairs:                                      # no pair may cross an enforced one or steal its partner
        for a, b in zip(range(0, 8), range(19, 11, -1)):
            assert not (i < a < j < b or a < i < b < j) and (i not in (a, b) or (i, j) == (a, b)) and (j not in (a, b) or (i, j) == (a, b))
