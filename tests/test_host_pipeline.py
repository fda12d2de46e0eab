"""Host side of the hot path (z/p, ScanFold-Fold aggregation, competition, writers) against golden files
written by the UNMODIFIED reference ScanFold.py (tests/golden/make_golden.py; its folds came from the CPU
oracle through the RNA shim).  The per-window engine outputs are replayed from trace.npz, the accumulators
come from a plain-loop numpy reference, so these tests need neither a GPU nor /root/reference.
Bar: every output file byte-identical."""
import json
import os
import types

import numpy as np
import pytest

from util import accumulate_numpy

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# (legacy_* holds the goldens of the legacy ScanFold-Scan.py surface: tests/test_legacy_scan.py)
CASES = sorted(d for d in os.listdir(GOLDEN)
               if os.path.exists(os.path.join(GOLDEN, d, "case.json")) and not d.startswith("legacy_"))
# written by the structure-extraction step that follows the hot path: compared separately below
# (and the full-length refold of --global_refold, which needs the fold engine)
NEXT_ROW_FILES = ("ExtractedStructures.gff3", "AllDBN-global_refold.txt")


def load_case(name):
    d = os.path.join(GOLDEN, name)
    case = json.load(open(os.path.join(d, "case.json")))
    lines = open(os.path.join(d, "input.fa")).read().split("\n")
    case["header"] = lines[0][1:]
    case["seq"] = lines[1].replace("T", "U")
    case["trace"] = np.load(os.path.join(d, "trace.npz"))
    case["dir"] = d
    return case


def replay_table(case):
    from scanfold_b200 import scan
    t = case["trace"]
    res = types.SimpleNamespace(W=case["W"], r=case["r"], n=case["n_windows"] + 1, mfe_dcal=t["mfe_dcal"],
                                native_unconstrained_dcal=t["native_unconstrained_dcal"], shuffle_dcal=t["shuffle_dcal"],
                                pair_tbl=t["pair_tbl"], centroid_tbl=t["centroid_tbl"], ed=t["ed"],
                                ensemble_dG=t["ensemble_dG"], ms_total=0.0, ms_mfe=0.0, n_launches=0)
    alln = t["alln"] if "alln" in t.files else np.zeros(case["n_windows"] + 1, dtype=bool)
    table = scan.table_from_result(res, 0, case["step"], case["n_windows"], True)
    if alln[:-1].any():
        table.alln = alln[:-1]                  # Q10 windows: no fold, no records
    if alln[-1]:
        table.final = None
    return table


def accumulate_inputs(table):
    """pair tables with the all-N windows marked (negative rows leave no records), z / MFE / ED in hundredths"""
    from scanfold_b200 import pipeline
    z100, mfe100, ed100 = pipeline.fold_inputs(table)
    pt = table.pair_tbl
    if table.alln is not None:
        pt = pt.copy()
        pt[table.alln] = -1
    return pt, z100, mfe100, ed100


def expected_files(case):
    exp = os.path.join(case["dir"], "expected")
    return {f: open(os.path.join(exp, f)).read() for f in os.listdir(exp)
            if f not in NEXT_ROW_FILES and "_motif_" not in f}


@pytest.mark.parametrize("name", CASES)
def test_outputs_byte_identical(name, tmp_path, monkeypatch):
    from scanfold_b200 import foldstep, pipeline
    case = load_case(name)
    table = replay_table(case)
    args = case["args"]
    stype = args[args.index("--type") + 1] if "--type" in args else "mono"
    chrom = args[args.index("--name") + 1] if "--name" in args else "UserInput"
    names = pipeline.RunNames(case["record"], case["header"].split()[0], case["W"], case["step"], case["r"], stype,
                              name=chrom)
    monkeypatch.chdir(tmp_path)
    temperature = int(args[args.index("-t") + 1]) if "-t" in args else 37
    minz = pipeline.write_scan_outputs(case["seq"], table, names, temperature, case["step"])
    comp = accumulate_numpy(case["L"], case["W"], case["step"], 0, *accumulate_inputs(table))
    ptable = foldstep.table_from_compact(*comp)
    competition = 0 if "-c" in args and args[args.index("-c") + 1] == "0" else 1
    pipeline.write_fold_outputs(case["seq"], ptable, names, minz, case["step"], by_ed="--by_ed" in args,
                                competition=competition, zscores=pipeline.zscore_total(table), filter_value=-2,
                                input_filename="input.fa")
    exp = expected_files(case)
    got = {f: open(os.path.join(tmp_path, f)).read() for f in os.listdir(tmp_path)}
    assert sorted(got) == sorted(exp)
    for f in sorted(exp):
        assert got[f] == exp[f], "%s differs in case %s" % (f, name)


def test_stats_match_golden_out_rows():
    """z / p columns of the .out file come out of stats.zscore_pvalue exactly (Appendix B Q1-Q3)"""
    from scanfold_b200 import stats
    for name in CASES:
        case = load_case(name)
        t = case["trace"]
        z, p = stats.zscore_pvalue(t["native_unconstrained_dcal"], t["shuffle_dcal"])
        out = [f for f in os.listdir(os.path.join(case["dir"], "expected")) if f.endswith(".out")][0]
        rows = open(os.path.join(case["dir"], "expected", out)).read().split("\n")[1:-1]
        assert len(rows) == case["n_windows"]
        alln = t["alln"] if "alln" in t.files else np.zeros(case["n_windows"] + 1, dtype=bool)
        for k, row in enumerate(rows):
            f = row.split("\t")
            if alln[k]:
                assert f[3:7] == ["0", "#DIV/0", "0", "0"] and f[8] == "." * 120, (name, k)      # Q10
                continue
            assert f[4] == str(float(z[k])) and f[5] == str(float(p[k])), (name, k)


def test_motif_extraction_vectors():
    """motifs.extract_motifs against known-answer vectors produced by EXECUTING the reference's own extraction block
    (ScanFold.py:1582-1716, see make_golden.py make_motif_vectors): pseudoknot characters, the lost opener at index 0,
    empty motifs from mismatched start / end lists, and the IndexError the reference raises when closers run out."""
    from scanfold_b200 import motifs
    vectors = json.load(open(os.path.join(GOLDEN, "motif_extract_vectors.json")))
    assert len(vectors) >= 30
    for v in vectors:
        if v["error"]:
            with pytest.raises(IndexError):
                motifs.extract_motifs(v["structure"] + "\n", v["seq"], log=lambda *a: None)
            continue
        got = motifs.extract_motifs(v["structure"] + "\n", v["seq"], log=lambda *a: None)
        assert [[m.i, m.j, m.sequence, m.structure] for m in got] == v["motifs"], v["structure"]


@pytest.mark.parametrize("name", CASES)
def test_motif_outputs_byte_identical(name, tmp_path, monkeypatch):
    """Structure-extraction outputs (motif .dbn / .ct, ExtractedStructures.gff3) of the host code against the golden
    files of the unmodified reference; the per-motif fold results come from the CPU oracle here (the checker), from
    the CUDA engine in tests/test_gpu_cli.py."""
    from oracle import oracle as O
    from scanfold_b200 import motifs, scan, stats
    case = load_case(name)
    if case.get("expect_fail"):
        pytest.skip("the reference dies before the structure-extraction step in this case (-c 0)")
    exp_dir = os.path.join(case["dir"], "expected")
    args = case["args"]
    chrom = args[args.index("--name") + 1] if "--name" in args else "UserInput"
    monkeypatch.chdir(tmp_path)
    found = motifs.extract_motifs(open(os.path.join(exp_dir, "Zavg_-2_pairs.dbn")).readlines()[2], case["seq"],
                                  log=lambda *a: None)
    results = []
    for m in found:
        sh = case["trace"]["motif_shuffles_%d" % m.number]
        bg_t = float(args[args.index("-t") + 1]) if "-t" in args else 37.0     # only energies() gets -t (ScanFold.py:1748)
        e, s = O.mfe(m.sequence, hc=m.structure)
        pf = O.pf(m.sequence, hc=m.structure)
        nat = O.mfe(m.sequence, structure=False, temperature=bg_t)[0]
        she = [O.mfe(bytes(row).decode(), structure=False, temperature=bg_t)[0] for row in sh]
        z, p = stats.zscore_pvalue(np.array([nat]), np.array([she]))
        results.append({"structure": s, "mfe": float(stats.round_energy([e])[0]), "z": float(z[0]), "p": float(p[0]),
                        "ed": float(scan.round_ed([pf["ed"]])[0])})
    motifs.write_motif_outputs(found, results, chrom, "ExtractedStructures.gff3")
    exp = {f: open(os.path.join(exp_dir, f)).read() for f in os.listdir(exp_dir)
           if f == "ExtractedStructures.gff3" or "_motif_" in f}
    got = {f: open(os.path.join(tmp_path, f)).read() for f in os.listdir(tmp_path)}
    assert sorted(got) == sorted(exp)
    for f in sorted(exp):
        assert got[f] == exp[f], "%s differs in case %s" % (f, name)
