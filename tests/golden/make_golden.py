#!/usr/bin/env python
"""Generates tests/golden/<case>/ by running the UNMODIFIED reference /root/reference/ScanFold.py.

The reference's fold engine (ViennaRNA) is absent, so `import RNA` resolves to tests/golden/stubs/RNA.py,
i.e. the Level-1 shim backed by the CPU oracle; Biopython is replaced by a 40-line FASTA reader and the
per-window process pools by an in-process map with a seeded `random` (stubs/sitecustomize.py).  Everything
else -- shuffling, z/p, pair records, aggregation, competition, every writer -- is the reference's own code.

Each case directory holds:
  case.json      command line, window geometry
  input.fa       (+ constraints.dbn / react.shape)
  trace.npz      per-window engine inputs/outputs recovered from the fold trace: the shuffled sequences the
                 reference drew (parity shuffles), energies, structures, ED, centroids; motif_shuffles_<n> = the
                 100 shuffles drawn for motif n by the structure-extraction step
  expected/      every output file the reference wrote (motif .ps placeholders dropped)

Run here (needs /root/reference):  python tests/golden/make_golden.py
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


def rand_seq(rng, n, alpha="ACGU"):
    return "".join(rng.choice(alpha) for _ in range(n))


def pt_from_db(s):
    pt = np.zeros(len(s), dtype=np.int16)
    stk = []
    for k, ch in enumerate(s):
        if ch == "(":
            stk.append(k)
        elif ch == ")":
            a = stk.pop()
            pt[a], pt[k] = k + 1, a + 1
    return pt


CASES = {
    # name: (L, seq seed, args, extras)
    "mono_w40": dict(L=150, seed=11, args=["-w", "40", "-r", "10"]),
    "mono_w60_gc": dict(L=260, seed=12, alpha="GGCCAU", args=["-w", "60", "-r", "12"]),
    "di_w40": dict(L=140, seed=13, args=["-w", "40", "-r", "10", "--type", "di"]),
    "step7_w40": dict(L=150, seed=14, args=["-w", "40", "-r", "10", "-s", "7"]),
    "hc_w40": dict(L=150, seed=15, args=["-w", "40", "-r", "10"], hc=True),
    "shape_w40": dict(L=150, seed=16, args=["-w", "40", "-r", "10", "--shapeD"], react=True),
    "motifs_w50": dict(L=0, seed=18, args=["-w", "50", "-r", "10"],
                       seq="AAUAC" + "GGGGCGCUUCGGCGCCCC" + "AUAAUUAAUA" + "GCCGGAUCGAAAGAUCCGGC" + "AAUAUAAUAAAUUA" +
                           "GGCACGGCUUUUGCCGUGCC" + "UUAUAAUAUA" + "CCGCGGAGAAAUCCGCGG" + "AUUAAUAUAAUAUUAAAUUAAUAAUAUUAAUA"),
    # C1 geometry (BASELINE.json configs[0]: 120-nt window, step 1, 100 mono shuffles) on a 400-nt record: pins the host
    # pipeline where the r = 100 z-score rounding and the exact-mean paths matter
    "c1_w120_r100": dict(L=400, seed=19, args=["-w", "120", "-r", "100"]),
    # f3 flag surface and quirks
    "by_ed_w40": dict(L=150, seed=21, args=["-w", "40", "-r", "10", "--by_ed"]),
    "c0_w40": dict(L=150, seed=22, args=["-w", "40", "-r", "10", "-c", "0"], expect_fail="FileNotFoundError"),
    "refold_w40": dict(L=170, seed=23, args=["-w", "40", "-r", "10", "--global_refold"]),
    "print_random_w40": dict(L=90, seed=24, args=["-w", "40", "-r", "6", "--print_random"], stdout=True),
    "t25_w40": dict(L=150, seed=27, args=["-w", "40", "-r", "10", "-t", "25"]),      # -t != 37: enthalpy-rescaled tables
    # the motif step folds at 37 C but scores against a background folded at -t (ScanFold.py:1733,1748)
    "motifs_t45_w50": dict(L=0, seed=28, args=["-w", "50", "-r", "10", "-t", "45"],
                           seq="AAUAC" + "GGGGCGCUUCGGCGCCCC" + "AUAAUUAAUA" + "GCCGGAUCGAAAGAUCCGGC" + "AAUAUAAUAAAUUA" +
                               "GGCACGGCUUUUGCCGUGCC" + "UUAUAAUAUA" + "CCGCGGAGAAAUCCGCGG" + "AUUAAUAUAAUAUUAAAUUAAUAAUAUUAAUA"),
    # Q10: runs of N -- 130 (every nucleotide still covered by partly-N windows) and 260 (22 nucleotides left without a record)
    "q10_n130_w120": dict(L=0, seed=25, args=["-w", "120", "-r", "6"], nrun=(150, 130, 140)),
    "q10_n260_w120": dict(L=0, seed=26, args=["-w", "120", "-r", "6"], nrun=(140, 260, 135)),
    "dna_name_w30": dict(L=100, seed=17, alpha="ACGT", args=["-w", "30", "-r", "8", "--name", "chrTest"],
                         header="rec17|extra|fields"),
}


def make_case(name, spec):
    rng = random.Random(spec["seed"])
    if spec.get("nrun"):
        a, b, c = spec["nrun"]
        spec = dict(spec, seq=rand_seq(rng, a) + "N" * b + rand_seq(rng, c))
    seq = spec.get("seq") or rand_seq(rng, spec["L"], spec.get("alpha", "ACGU"))
    header = spec.get("header", name)
    work = tempfile.mkdtemp(prefix="golden_")
    fasta = os.path.join(work, "input.fa")
    with open(fasta, "w") as f:
        f.write(">%s\n%s\n" % (header, seq))
    args = list(spec["args"])
    aux = {}
    if spec.get("hc"):
        hc = "".join(rng.choice("......x") for _ in seq)
        path = os.path.join(work, "constraints.dbn")
        with open(path, "w") as f:
            f.write(">%s\n%s\n%s\n" % (name, seq, hc))
        args += ["--constraints", path]          # must be absolute (Appendix B Q8)
        aux["constraints.dbn"] = path
    if spec.get("react"):
        nrng = np.random.default_rng(spec["seed"])
        vals = np.clip(nrng.exponential(0.4, len(seq)), 0, 4)
        path = os.path.join(work, "react.shape")
        with open(path, "w") as f:
            for k, v in enumerate(vals):
                if k % 23 == 5:
                    continue                      # gap -> parser fills -999
                f.write("%d\t%s\t%s\n" % (k + 1, seq[k], "NA" if k % 17 == 3 else "%.4f" % v))
        args += ["--react", "react.shape"]       # resolved against the original cwd
        aux["react.shape"] = path
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "stubs"), REF])
    env["SCANFOLD_GOLDEN_SEED"] = str(1000 + spec["seed"])
    trace_path = os.path.join(work, "trace.json")
    env["SCANFOLD_GOLDEN_TRACE"] = trace_path
    cmd = [sys.executable, "-W", "ignore", os.path.join(REF, "ScanFold.py"), "input.fa"] + args
    p = subprocess.run(cmd, cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if spec.get("expect_fail"):
        if p.returncode == 0 or spec["expect_fail"] not in p.stdout:
            print(p.stdout[-3000:])
            raise SystemExit("reference run was expected to die with %s for %s" % (spec["expect_fail"], name))
    elif p.returncode != 0:
        print(p.stdout[-3000:])
        raise SystemExit("reference run failed for " + name)
    rec = header.split("|")[0]
    outdir = os.path.join(work, rec)

    # ---- rebuild per-window arrays from the fold trace
    trace = json.load(open(trace_path))
    W = int(args[args.index("-w") + 1])
    r = int(args[args.index("-r") + 1])
    step = int(args[args.index("-s") + 1]) if "-s" in args else 1
    L = len(seq)
    spec = dict(spec, L=L)
    nwin = (L - W) // step + 1
    n = nwin + 1
    shuf = np.zeros((n, r, W), dtype=np.uint8)
    mfe = np.zeros(n, dtype=np.int32)
    nat = np.zeros(n, dtype=np.int32)
    she = np.zeros((n, r), dtype=np.int32)
    pair_tbl = np.zeros((n, W), dtype=np.int16)
    cen_tbl = np.zeros((n, W), dtype=np.int16)
    ed = np.zeros(n)
    dG = np.zeros(n)
    pos = 0
    alln = np.zeros(n, dtype=bool)
    for w in range(n):
        start = w * step if w < nwin else L - W
        frag = seq[start:start + W]
        if (w < nwin and W == 120 and frag == "N" * 120) or (w == nwin and frag == "N" * W):
            alln[w] = True                       # Q10: the reference never folds this window
            continue
        grp = trace[pos:pos + r + 3]
        pos += r + 3
        ops = [g["op"] for g in grp]
        first_mfe = ops.index("mfe")
        first_pf = ops.index("pf")
        assert sorted(ops[:2]) == ["mfe", "pf"] and ops[2:] == ["mfe"] * (r + 1), (name, w, ops)
        m, q = grp[first_mfe], grp[first_pf]
        mfe[w], pair_tbl[w] = m["e"], pt_from_db(m["s"])
        ed[w], dG[w], cen_tbl[w] = q["ed"], q["dG"], pt_from_db(q["centroid"])
        nat[w] = grp[2]["e"]
        for k in range(r):
            she[w, k] = grp[3 + k]["e"]
            shuf[w, k] = np.frombuffer(grp[3 + k]["seq"].encode(), dtype=np.uint8)
    if "--global_refold" in args:            # three full-length folds (ScanFold.py:1518-1539)
        assert [g["op"] for g in trace[pos:pos + 3]] == ["mfe"] * 3, (name, "refold trace")
        pos += 3
    if spec.get("expect_fail"):
        assert pos == len(trace), (name, "ops after the failure point")
    # ---- motif step (ScanFold.py:1724-1750): per motif pf (hc), pf (RNA.pf_fold), mfe (hc), then native + 100 shuffles
    motif_shuffles = {}
    k_motif = 0
    while pos < len(trace):
        grp = trace[pos:pos + 104]
        pos += 104
        assert [g["op"] for g in grp] == ["pf", "pf", "mfe"] + ["mfe"] * 101, (name, "motif trace")
        k_motif += 1
        motif_shuffles["motif_shuffles_%d" % k_motif] = np.stack(
            [np.frombuffer(g["seq"].encode(), dtype=np.uint8) for g in grp[4:104]])
    dst = os.path.join(HERE, name)
    if os.path.exists(dst):
        shutil.rmtree(dst)
    os.makedirs(os.path.join(dst, "expected"))
    shutil.copy(fasta, os.path.join(dst, "input.fa"))
    for fn, path in aux.items():
        shutil.copy(path, os.path.join(dst, fn))
    np.savez_compressed(os.path.join(dst, "trace.npz"), shuffles=shuf, mfe_dcal=mfe, native_unconstrained_dcal=nat,
                        shuffle_dcal=she, pair_tbl=pair_tbl, centroid_tbl=cen_tbl, ed=ed, ensemble_dG=dG, alln=alln, **motif_shuffles)
    kept = []
    for fn in sorted(os.listdir(outdir)):
        if fn.endswith(".ps"):
            continue
        shutil.copy(os.path.join(outdir, fn), os.path.join(dst, "expected", fn))
        kept.append(fn)
    rel_args = [("constraints.dbn" if a == aux.get("constraints.dbn") else a) for a in args]
    extra = {}
    if spec.get("stdout"):                   # what the reference printed per window (--print_random energy lists)
        extra["stdout_lists"] = [ln for ln in p.stdout.split("\n") if ln.startswith("[")]
    if spec.get("expect_fail"):
        extra["expect_fail"] = spec["expect_fail"]
    json.dump({"name": name, "args": rel_args, "record": rec, "L": L, "W": W, "step": step, "r": r,
               "n_windows": nwin, "files": kept, "python": sys.version.split()[0], **extra},
              open(os.path.join(dst, "case.json"), "w"), indent=1)
    shutil.rmtree(work)
    print("%-14s windows=%d files=%d" % (name, nwin, len(kept)))


def make_motif_vectors():
    """Known-answer vectors for motifs.extract_motifs: the reference's own extraction block (ScanFold.py, from
    `structure_raw = filter2constraints` to the ExtractedStructure list) executed on synthetic dot-bracket lines,
    including pseudoknot characters, an opener at index 0 and nested helices."""
    import textwrap
    import types
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    sys.path.insert(0, REF)
    import ScanFoldFunctions as SFF
    src = open(os.path.join(REF, "ScanFold.py")).read().split("\n")
    a = next(k for k, ln in enumerate(src) if ln.strip() == "structure_raw = filter2constraints")
    b = next(k for k, ln in enumerate(src) if "extracted_structure_list.append(" in ln)
    block = textwrap.dedent("\n".join(src[a:b + 2]))
    rng = random.Random(77)
    lines = ["..((((....))))..((...))..", "((((....))))..((...))", "..((..((...))..((...))..))..(((...)))",
             "..((..<<..))..>>..((...))..", "..<<<...>>>..{{..}}..((...))", ".((.{.)).}.((...)).", "....", "",
             "((...))", ".(((...)))(((...))).", "..((..<..))..>..", "..(((...)))<<<...>>>"]
    for _ in range(20):          # random balanced structures with a few pseudoknot characters sprinkled in
        n = rng.randint(20, 70)
        st, depth = [], 0
        for k in range(n):
            c = rng.choice("..((()))..<>{}" if k % 7 == 3 else "...(())")
            if c == ")" and depth == 0:
                c = "."
            if c == "(":
                depth += 1
            if c == ")":
                depth -= 1
            st.append(c)
        st += [")"] * depth
        lines.append("".join(st))
    vectors = []
    for st in lines:
        seq = "".join(rng.choice("ACGU") for _ in st) + "ACG"
        nuc_dict = {k + 1: SFF.NucZscore(ch, k + 1) for k, ch in enumerate(seq)}
        ns = {"filter2constraints": st + "\n", "seq": seq, "cur_record": types.SimpleNamespace(seq=seq),
              "nuc_dict": nuc_dict, "NucStructure": SFF.NucStructure, "ExtractedStructure": SFF.ExtractedStructure,
              "bond_order": [], "bond_count": 0, "print": lambda *a, **k: None}
        try:
            exec(block, ns)
            out = [[es.i, es.j, es.sequence, es.structure] for es in ns["extracted_structure_list"]]
            err = None
        except Exception as ex:          # unbalanced openers / closers crash the reference (IndexError)
            out, err = None, type(ex).__name__
        vectors.append({"structure": st, "seq": seq, "motifs": out, "error": err})
    json.dump(vectors, open(os.path.join(HERE, "motif_extract_vectors.json"), "w"), indent=0)
    print("motif extraction vectors: %d (%d raise)" % (len(vectors), sum(v["error"] is not None for v in vectors)))


if __name__ == "__main__":
    sel = sys.argv[1:] or sorted(CASES) + ["motif_vectors"]
    for nm in sel:
        if nm == "motif_vectors":
            make_motif_vectors()
        else:
            make_case(nm, CASES[nm])
