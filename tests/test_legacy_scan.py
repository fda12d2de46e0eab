"""The legacy scan surface: scanfold_b200.legacy_scan against the `.txt` tables and the screen
output of the UNMODIFIED /root/reference/ScanFold-Scan.py (tests/golden/legacy_*, made by make_golden_legacy.py).

CPU: header / row formatting and the legacy z / p arithmetic on the fold results recorded in the golden trace.
GPU: the command line end to end through the C-ABI, fed the shuffles the reference drew -- byte-identical file and
screen output."""
import io
import json
import os
import shutil

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(d for d in os.listdir(GOLDEN) if d.startswith("legacy_"))


def load(name):
    d = os.path.join(GOLDEN, name)
    case = json.load(open(os.path.join(d, "case.json")))
    seq = open(os.path.join(d, "input.fa")).read().split("\n")[1]
    tr = np.load(os.path.join(d, "trace.npz"))
    hc_line = None
    if os.path.exists(os.path.join(d, "constraints.dbn")):
        hc_line = list(open(os.path.join(d, "constraints.dbn")).readlines()[2])
    return d, case, seq, tr, hc_line


def table_from_trace(case, tr):
    from scanfold_b200 import scan, stats
    t = scan.WindowTable()
    t.W, t.step, t.r = case["W"], case["step"], case["r"]
    t.start1 = np.arange(case["n_windows"], dtype=np.int64) * case["step"] + 1
    t.end1 = t.start1 + case["W"] - 1
    t.mfe_dcal = tr["mfe_dcal"]
    t.mfe = stats.round_energy(tr["mfe_dcal"])
    t.ed = scan.round_ed(tr["ed"])
    t.native_unconstrained_dcal = tr["native_unconstrained_dcal"]
    t.shuffle_dcal = tr["shuffle_dcal"]
    t.pair_tbl, t.centroid_tbl = tr["pair_tbl"], tr["centroid_tbl"]
    return t


def test_cases_exist():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_rows_and_screen_output_from_trace(name):
    from scanfold_b200 import legacy_scan
    d, case, seq, tr, hc_line = load(name)
    args = legacy_scan.build_parser().parse_args(["-i", "input.fa"] + case["args"])
    assert legacy_scan.output_name("input.fa", int(args.w), int(args.s), int(args.r), args.type) == case["output"]
    out = io.StringIO()
    text = legacy_scan.format_record(name, seq, case["W"], case["step"], case["r"], int(args.t), table_from_trace(case, tr),
                                     hc_line=hc_line, print_to_screen=args.print_to_screen,
                                     print_random=str(args.print_random), out=out)
    assert text == open(os.path.join(d, "expected", case["output"])).read()
    # what the script printed per window (-p blocks, --print_random lists); the record banner is main()'s
    banner = "Scanning sequence %s\nSequence Length: %dnt long.\n" % (name, len(seq))
    if hc_line is not None:
        banner += "Considering constraint input\nConstraint list is %dnt long.\n" % (len(hc_line) - 1)
    assert banner + out.getvalue() == open(os.path.join(d, "stdout.txt")).read()


def test_legacy_defaults_and_flag_names():
    from scanfold_b200 import legacy_scan
    a = legacy_scan.build_parser().parse_args(["-i", "x.fa"])
    assert (a.s, a.w, a.r, a.t, a.type, a.print_to_screen, a.print_random, a.constraints) == (10, 120, 50, 37, "mono", False, "off", None)
    assert legacy_scan.window_starts(150, 40, 10) == list(range(0, 111, 10))


def test_legacy_stats_quirks():
    """population standard deviation over native + r shuffles, mean over the first r - 1 shuffles, strict '<'"""
    from scanfold_b200 import legacy_scan
    nat, sh = -500, np.array([-300, -700, -400, -100], dtype=np.int32)
    e, z, p = legacy_scan.legacy_stats(nat, sh, 4)
    assert e[0] == float(np.float32(-5.0)) and len(e) == 5
    want = (e[0] - np.mean(e[1:4])) / np.std(e)
    assert z == round(want, 2) and p == round(1 / 5, 2)
    _, z0, p0 = legacy_scan.legacy_stats(0, np.zeros(4, dtype=np.int32), 4)
    assert z0 == "#DIV/0!" and p0 == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_legacy_cli_byte_identical_on_gpu(name, tmp_path, monkeypatch, engine, capsys):
    from scanfold_b200 import legacy_scan
    d, case, seq, tr, hc_line = load(name)
    for f in ("input.fa", "constraints.dbn", "trace.npz"):
        if os.path.exists(os.path.join(d, f)):
            shutil.copy(os.path.join(d, f), tmp_path / f)
    monkeypatch.chdir(tmp_path)
    legacy_scan.main(["-i", "input.fa"] + case["args"] + ["--parity_shuffles", "trace.npz"])
    assert open(tmp_path / case["output"]).read() == open(os.path.join(d, "expected", case["output"])).read()
    assert capsys.readouterr().out == open(os.path.join(d, "stdout.txt")).read()
