"""End-to-end drop-in test on the GPU: the ScanFold.py-compatible CLI, fed the shuffles the reference drew
(parity mode), must write files byte-identical to the golden outputs of the unmodified reference."""
import json
import os
import shutil

import numpy as np
import pytest

from test_host_pipeline import CASES, GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_cli_outputs_byte_identical(name, tmp_path, monkeypatch, engine, capsys):
    from scanfold_b200 import cli
    d = os.path.join(GOLDEN, name)
    case = json.load(open(os.path.join(d, "case.json")))
    for f in ("input.fa", "constraints.dbn", "react.shape", "trace.npz"):
        if os.path.exists(os.path.join(d, f)):
            shutil.copy(os.path.join(d, f), tmp_path / f)
    args = [a if a != "constraints.dbn" else str(tmp_path / "constraints.dbn") for a in case["args"]]
    monkeypatch.chdir(tmp_path)
    argv = ["input.fa"] + args + ["--parity_shuffles", "trace.npz"]
    if case.get("expect_fail"):             # -c 0: the reference writes the .dp files, then dies opening a dbn file
        with pytest.raises(FileNotFoundError):
            cli.main(argv)
    else:
        cli.main(argv)
    if case.get("stdout_lists"):            # --print_random: the energy list of every window, as the reference prints it
        printed = [ln for ln in capsys.readouterr().out.split("\n") if ln.startswith("[")]
        assert printed == case["stdout_lists"]
    out = tmp_path / case["record"]
    exp_dir = os.path.join(d, "expected")
    exp = sorted(os.listdir(exp_dir))          # scan, ScanFold-Fold and structure-extraction outputs
    assert sorted(os.listdir(out)) == exp
    for f in exp:
        assert open(out / f).read() == open(os.path.join(exp_dir, f)).read(), "%s differs in case %s" % (f, name)


def test_accumulator_matches_numpy_and_merges(engine):
    """sfb_accumulate_*: exact sums vs the plain-loop reference; two shards + halo merge == one shard"""
    import torch
    from util import accumulate_numpy
    from scanfold_b200 import foldstep
    rng = np.random.default_rng(2)
    L, W, step = 300, 40, 3
    nwin = (L - W) // step + 1
    pt = np.zeros((nwin, W), dtype=np.int16)
    for w in range(nwin):
        for _ in range(6):
            i = int(rng.integers(0, W - 5))
            j = int(rng.integers(i + 4, W))
            if pt[w, i] == 0 and pt[w, j] == 0:
                pt[w, i], pt[w, j] = j + 1, i + 1
    z = rng.integers(-400, 300, nwin).astype(np.int32)
    m = rng.integers(-3000, 0, nwin).astype(np.int32)
    e = rng.integers(0, 4000, nwin).astype(np.int32)
    ref = accumulate_numpy(L, W, step, 0, pt, z, m, e)
    acc = engine.Accumulator(L, W, step, 0, pt, z, m, e)
    got = acc.compact()
    for a, b in zip(got, ref):
        assert np.array_equal(np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64))
    # shard the windows in two, move the overlap rows of shard A into shard B through device buffers
    h = nwin // 2
    A = engine.Accumulator(L, W, step, 0, pt[:h], z[:h], m[:h], e[:h])
    B = engine.Accumulator(L, W, step, h, pt[h:], z[h:], m[h:], e[h:])
    own_a = B.nt0 - A.nt0                     # rows of A before B's first nucleotide
    halo = A.n_nt - own_a
    ncol = 2 * W - 1
    cnt = torch.zeros(halo * ncol, dtype=torch.int32, device="cuda")
    fst = torch.zeros(halo * ncol, dtype=torch.int32, device="cuda")
    sums = torch.zeros(6 * halo * ncol, dtype=torch.int64, device="cuda")
    A.export_rows(own_a, halo, cnt.data_ptr(), fst.data_ptr(), sums.data_ptr())
    B.merge_rows(0, halo, cnt.data_ptr(), fst.data_ptr(), sums.data_ptr())
    ta = foldstep.table_from_compact(*A.compact(0, own_a))
    tb = foldstep.table_from_compact(*B.compact())
    whole = foldstep.table_from_compact(*got)
    both = foldstep.concat_tables([ta, tb])
    for name in ("nt_ptr", "partner", "count", "first_seen", "sums"):
        assert np.array_equal(getattr(both, name), getattr(whole, name)), name
    for x in (acc, A, B):
        x.close()


def test_repeated_plans_reuse_device_memory(engine):
    """The library keeps freed device blocks for the next plan (DevPool in api.cu): alternating plans of different
    geometry must give the same results as the first time each one ran."""
    rng = np.random.default_rng(11)
    seq_a = "".join("ACGU"[k] for k in rng.integers(0, 4, 400))
    seq_b = "".join("ACGU"[k] for k in rng.integers(0, 4, 900))

    def run(seq, W, r):
        res = engine.scan(seq, W, 1, r, seed=5)
        return res.mfe_dcal.copy(), res.shuffle_dcal.copy(), res.ed.copy(), res.pair_tbl.copy()

    first_a, first_b = run(seq_a, 60, 8), run(seq_b, 120, 5)
    for _ in range(2):
        for first, args in ((first_a, (seq_a, 60, 8)), (first_b, (seq_b, 120, 5))):
            again = run(*args)
            for x, y in zip(first, again):
                assert np.array_equal(x, y)


def test_cli_two_records_device_shuffles(tmp_path, monkeypatch, engine):
    """Stock run (device shuffles, no parity file) on a two-record FASTA: one output folder per record with the full
    file set, and the structure-extraction outputs are consistent with each other."""
    from scanfold_b200 import cli
    rng = np.random.default_rng(3)
    hp = "GGGGCGCUUCGGCGCCCC"
    recs = {"recA": "AAUAC" + hp + "AUAAUUAAUAUAUUAAUUA" + "GCCGGAUCGAAAGAUCCGGC" + "AAUAUAAUAAAUUAUAUA",
            "recB": "".join("ACGU"[k] for k in rng.integers(0, 4, 130))}
    with open(tmp_path / "two.fa", "w") as f:
        for k, v in recs.items():
            f.write(">%s\n%s\n" % (k, v))
    monkeypatch.chdir(tmp_path)
    cli.main(["two.fa", "-w", "40", "-r", "20", "--seed", "9"])
    for name, seq in recs.items():
        out = tmp_path / name
        files = sorted(os.listdir(out))
        assert "ExtractedStructures.gff3" in files and "AllDBN.txt" in files and "Zavg_-2_pairs.ct" in files
        assert any(f.endswith(".out") for f in files) and any(f.endswith(".scan-zscores.wig") for f in files)
        gff = [ln for ln in open(out / "ExtractedStructures.gff3").read().split("\n") if ln]
        motifs = sorted(f for f in files if "_motif_" in f)
        assert len(motifs) == 2 * len(gff)          # one .dbn and one .ct per gff3 line
        for k, ln in enumerate(gff, 1):
            fields = ln.split("\t")
            start, end = int(fields[3]), int(fields[4])
            attrs = dict(a.split("=", 1) for a in fields[8].split(";")[1:])
            assert attrs["sequence"] == seq[start - 1:end]
            dbn = open(out / ("UserInput_motif_%d.dbn" % k)).read().split("\n")
            assert dbn[1] == attrs["sequence"] and dbn[2] == attrs["refoldedMFE"]
            assert len(attrs["structure"]) == len(attrs["sequence"])
    assert any(ln for ln in open(tmp_path / "recA" / "ExtractedStructures.gff3"))   # the designed hairpins are found
