"""Multi-rank host logic on CPU: two gloo ranks shard the windows of one golden case, exchange the accumulator
halo, and rank 0 must end up with exactly the single-process partner table (and byte-identical output files)."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    from util import NumpyAccumulator, accumulate_numpy
    from test_host_pipeline import load_case, replay_table
    from scanfold_b200 import foldstep, multigpu, pipeline
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case = load_case(%(case)r)
    table = replay_table(case)
    z100, mfe100, ed100 = pipeline.fold_inputs(table)
    W, step, n = case["W"], case["step"], case["n_windows"]
    w0, w1 = multigpu.shard_windows(n, world, rank)
    acc = NumpyAccumulator(case["L"], W, step, w0, table.pair_tbl[w0:w1], z100[w0:w1], mfe100[w0:w1], ed100[w0:w1]) \
        if w1 > w0 else None
    own_table = multigpu.own_partner_table(acc, W, step, rank, world, dist, n)
    agg, log_text, cnt_text = multigpu.aggregate_distributed(own_table, case["seq"], rank, world, dist, with_logs=True)
    digest = multigpu.table_checksum(own_table, rank, world, dist)
    merged = multigpu.gather_tables(own_table, rank, world, dist)
    # the per-window columns travel as flat byte tensors too (no pickled objects)
    from scanfold_b200 import scan
    shard = scan.empty_table(W, step, table.r, w0)
    for k in ("start1", "end1", "mfe_dcal", "mfe", "z", "p", "ed", "pair_tbl", "centroid_tbl",
              "native_unconstrained_dcal", "shuffle_dcal"):
        setattr(shard, k, getattr(table, k)[w0:w1])
    shard.final = table.final if w1 == n and w1 > w0 else None
    whole_w = multigpu.gather_window_tables(shard, rank, world, dist, with_shuffle_energies=True)
    if rank == 0:
        for k in ("start1", "mfe", "z", "p", "ed", "pair_tbl", "centroid_tbl", "shuffle_dcal"):
            assert np.array_equal(getattr(whole_w, k), getattr(table, k)[:n]), k
        assert whole_w.final == table.final
        whole = foldstep.table_from_compact(*accumulate_numpy(case["L"], W, step, 0, table.pair_tbl, z100, mfe100, ed100))
        for name in ("nt_ptr", "coord", "partner", "count", "first_seen", "sums"):
            assert np.array_equal(getattr(merged, name), getattr(whole, name)), name
        # every rank aggregated its own nucleotides: results and log text equal the single-process aggregation
        import io
        lt, st = io.StringIO(), io.StringIO()
        ref = foldstep.aggregate(whole, case["seq"], lt, st)
        for name in ("coord", "part", "cov_z", "mean_z", "mean_mfe", "mean_ed", "total_windows", "num_bp"):
            assert np.array_equal(getattr(agg, name), getattr(ref, name)), name
        assert log_text == lt.getvalue() and cnt_text == st.getvalue()
        assert digest == multigpu.table_checksum(whole, 0, 1, None)
        print("MERGE_OK", len(merged.partner))
    dist.destroy_process_group()
""")


# world 4 / 8 on the 111-window case: shards shorter than the W - step overlap, so halo rows cross several ranks
# (the single-hop exchange lost them); world 8 on step7_w40 has 16 windows: 2 per rank
@pytest.mark.parametrize("case,world", [("mono_w40", 2), ("step7_w40", 2), ("mono_w60_gc", 3), ("mono_w40", 4),
                                        ("step7_w40", 8)])
def test_two_rank_halo_exchange_equals_single_process(case, world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "case": case})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 400), str(script)]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "MERGE_OK" in p.stdout, p.stdout[-3000:]


def test_shard_windows_partition():
    from scanfold_b200 import multigpu
    for total in (1, 7, 100, 29784):
        for world in (1, 2, 3, 8):
            edges = [multigpu.shard_windows(total, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
