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
def test_cli_outputs_byte_identical(name, tmp_path, monkeypatch, engine):
    from scanfold_b200 import cli
    d = os.path.join(GOLDEN, name)
    case = json.load(open(os.path.join(d, "case.json")))
    for f in ("input.fa", "constraints.dbn", "react.shape", "trace.npz"):
        if os.path.exists(os.path.join(d, f)):
            shutil.copy(os.path.join(d, f), tmp_path / f)
    args = [a if a != "constraints.dbn" else str(tmp_path / "constraints.dbn") for a in case["args"]]
    monkeypatch.chdir(tmp_path)
    cli.main(["input.fa"] + args + ["--parity_shuffles", "trace.npz"])
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
