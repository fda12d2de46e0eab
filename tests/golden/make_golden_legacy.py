#!/usr/bin/env python
"""Generates tests/golden/legacy_<case>/ by running the UNMODIFIED legacy script /root/reference/ScanFold-Scan.py.

Same harness as make_golden.py: `import RNA` resolves to the Level-1 shim backed by the CPU oracle, Biopython to the
FASTA stub, the process pools to an in-process map with a seeded `random` (stubs/).  Each case directory holds
  case.json      command line, window geometry
  input.fa       (+ constraints.dbn)
  trace.npz      per-window arrays recovered from the fold trace: the shuffles the script drew (parity shuffles),
                 energies, structures, ED, centroids
  expected/      the `.txt` table the script wrote;  stdout.txt = everything it printed
Run here (needs /root/reference):  python tests/golden/make_golden_legacy.py
The fixtures are oracle-derived, NOT ViennaRNA-derived (parity unpinned, see oracle/sf_oracle.h).
"""
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
from make_golden import pt_from_db, rand_seq  # noqa: E402

CASES = {
    "legacy_mono_w40": dict(L=150, seed=31, args=["-w", "40", "-s", "10", "-r", "10"]),
    "legacy_di_w30_s7": dict(L=120, seed=32, args=["-w", "30", "-s", "7", "-r", "8", "-type", "di", "--print_random", "on"]),
    "legacy_hc_t25_w40": dict(L=130, seed=33, args=["-w", "40", "-s", "5", "-r", "9", "-t", "25", "-p"], hc=True),
    # a 120-nt window that is all N (never folded) between windows that are partly N, and a poly-A stretch whose
    # energies are all zero (standard deviation 0: "#DIV/0!")
    "legacy_n_w120": dict(L=0, seed=34, args=["-w", "120", "-s", "10", "-r", "6"], parts=(("rand", 130), ("N", 140), ("A", 150))),
    "legacy_defaults": dict(L=260, seed=35, args=[]),      # -s 10 -w 120 -r 50
}


def make_case(name, spec):
    rng = random.Random(spec["seed"])
    if spec.get("parts"):
        seq = "".join(rand_seq(rng, n) if kind == "rand" else kind * n for kind, n in spec["parts"])
    else:
        seq = rand_seq(rng, spec["L"], spec.get("alpha", "ACGU"))
    work = tempfile.mkdtemp(prefix="golden_legacy_")
    with open(os.path.join(work, "input.fa"), "w") as f:
        f.write(">%s\n%s\n" % (name, seq))
    args = list(spec["args"])
    if spec.get("hc"):
        hc = "".join(rng.choice("......x<>") for _ in seq)
        with open(os.path.join(work, "constraints.dbn"), "w") as f:
            f.write(">%s\n%s\n%s\n" % (name, seq, hc))
        args += ["-c", "constraints.dbn"]
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "stubs"), REF])
    env["SCANFOLD_GOLDEN_SEED"] = str(1000 + spec["seed"])
    trace_path = os.path.join(work, "trace.json")
    env["SCANFOLD_GOLDEN_TRACE"] = trace_path
    cmd = [sys.executable, "-W", "ignore", os.path.join(REF, "ScanFold-Scan.py"), "-i", "input.fa"] + args
    p = subprocess.run(cmd, cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if p.returncode != 0:
        print(p.stdout[-2000:], p.stderr[-3000:])
        raise SystemExit("legacy reference run failed for " + name)
    W = int(args[args.index("-w") + 1]) if "-w" in args else 120
    r = int(args[args.index("-r") + 1]) if "-r" in args else 50
    step = int(args[args.index("-s") + 1]) if "-s" in args else 10
    stype = args[args.index("-type") + 1] if "-type" in args else "mono"
    L = len(seq)
    starts = list(range(0, L - W + 1, step))
    n = len(starts)
    trace = json.load(open(trace_path))
    shuf = np.zeros((n, r, W), dtype=np.uint8)
    mfe = np.zeros(n, dtype=np.int32)
    nat = np.zeros(n, dtype=np.int32)
    she = np.zeros((n, r), dtype=np.int32)
    pair_tbl = np.zeros((n, W), dtype=np.int16)
    cen_tbl = np.zeros((n, W), dtype=np.int16)
    ed = np.zeros(n)
    pos = 0
    for k, i in enumerate(starts):
        frag = seq[i:i + W]
        if frag == "N" * 120:
            shuf[k] = ord("N")
            continue
        grp = trace[pos:pos + r + 6]
        pos += r + 6
        # pf, pf (RNA.pf_fold), mfe, then mfe + pf again (under the constraint line if one was given), then native + r shuffles
        assert [g["op"] for g in grp] == ["pf", "pf", "mfe", "mfe", "pf"] + ["mfe"] * (r + 1), (name, k)
        m, q = grp[3], grp[4]
        mfe[k], pair_tbl[k] = m["e"], pt_from_db(m["s"])
        ed[k], cen_tbl[k] = q["ed"], pt_from_db(q["centroid"])
        nat[k] = grp[5]["e"]
        for z in range(r):
            she[k, z] = grp[6 + z]["e"]
            shuf[k, z] = np.frombuffer(grp[6 + z]["seq"].encode(), dtype=np.uint8)
    assert pos == len(trace), (name, "unexpected folds", pos, len(trace))
    dst = os.path.join(HERE, name)
    if os.path.exists(dst):
        shutil.rmtree(dst)
    os.makedirs(os.path.join(dst, "expected"))
    shutil.copy(os.path.join(work, "input.fa"), os.path.join(dst, "input.fa"))
    if spec.get("hc"):
        shutil.copy(os.path.join(work, "constraints.dbn"), os.path.join(dst, "constraints.dbn"))
    np.savez_compressed(os.path.join(dst, "trace.npz"), shuffles=shuf, mfe_dcal=mfe, native_unconstrained_dcal=nat,
                        shuffle_dcal=she, pair_tbl=pair_tbl, centroid_tbl=cen_tbl, ed=ed)
    out_name = "input.fa.forward.win_%d.stp_%d.rnd_%d.shfl_%s.txt" % (W, step, r, stype)
    shutil.copy(os.path.join(work, out_name), os.path.join(dst, "expected", out_name))
    with open(os.path.join(dst, "stdout.txt"), "w") as f:
        f.write(p.stdout)
    json.dump({"name": name, "args": args, "L": L, "W": W, "step": step, "r": r, "type": stype, "n_windows": n,
               "output": out_name, "python": sys.version.split()[0], "numpy": np.__version__},
              open(os.path.join(dst, "case.json"), "w"), indent=1)
    shutil.rmtree(work)
    print("%-20s windows=%d" % (name, n))


if __name__ == "__main__":
    for nm in sys.argv[1:] or sorted(CASES):
        make_case(nm, CASES[nm])
