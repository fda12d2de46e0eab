#!/usr/bin/env python3
"""Generate scanfold_b200/params/rna_turner2004_besteffort.par.

Why this exists
---------------
The reference (moss-lab/ScanFold) gets every energy from ViennaRNA's built-in
Turner-2004 tables (``RNA.md()`` at ScanFold.py:212, ``RNA.fold_compound`` at
ScanFold.py:494, ScanFoldFunctions.py:786).  ViennaRNA is NOT available in this
environment (no wheel, no source, no network) and no file on the box contains
``rna_turner2004.par``.  Both engines in this repo (CPU oracle and CUDA) are
fully table driven and load ViennaRNA's "## RNAfold parameter file v2.0" text
format, so dropping the real ``rna_turner2004.par`` in (``--params`` /
``SCANFOLD_PARAMS``) gives the real model.  Until then this script writes a
best-effort stand-in with the same shape:

* RECALLED blocks   - values written down from memory of the published
  Turner-2004 / ViennaRNA 2.x file (stack, hairpin/bulge/interior initiation,
  ML, NINIO, Misc, dangles, terminal mismatches, special hairpins).  They are
  believed close, but are unverified => "parity unpinned".
* SYNTHESISED blocks - int11 / int21 / int22 are produced by a documented
  rule (initiation + AU/GU closure penalty + mismatch bonus) because ~11k
  tabulated integers cannot be recalled.  They are NOT the Turner values.
* Enthalpy blocks are written as copies of the 37C free energies scaled by a
  fixed factor purely so that the file parses; folding at T != 37 with this
  stand-in is refused by the loaders (flag ``besteffort`` in the header).

Run:  python tools/gen_besteffort_par.py
"""
import os

PAIRS = ["CG", "GC", "GU", "UG", "AU", "UA", "NS"]
NUC = ["N", "A", "C", "G", "U"]
INF = 10000000

stack = [
    [-240, -330, -210, -140, -210, -210, -140],
    [-330, -340, -250, -150, -220, -240, -150],
    [-210, -250, 130, -50, -140, -130, 130],
    [-140, -150, -50, 30, -60, -100, 30],
    [-210, -220, -140, -60, -110, -90, -60],
    [-210, -240, -130, -100, -90, -130, -90],
    [-140, -150, 130, 30, -60, -90, 130],
]
stack_dH = [
    [-1060, -1340, -1210, -560, -1050, -1040, -560],
    [-1340, -1490, -1260, -830, -1140, -1240, -830],
    [-1210, -1260, -1460, -1350, -880, -1280, -880],
    [-560, -830, -1350, -930, -320, -700, -320],
    [-1050, -1140, -880, -320, -940, -680, -320],
    [-1040, -1240, -1280, -700, -680, -770, -680],
    [-560, -830, -880, -320, -320, -680, -320],
]

# 7 x (5x5); rows = 5' neighbour (N,A,C,G,U), cols = 3' neighbour
mismatch_hairpin = [
    [[-80, -100, -110, -100, -80], [-140, -150, -150, -140, -150], [-80, -100, -110, -100, -80],
     [-150, -230, -150, -240, -150], [-100, -100, -140, -100, -210]],
    [[-50, -110, -70, -110, -50], [-110, -110, -150, -130, -150], [-50, -110, -70, -110, -50],
     [-150, -250, -150, -220, -150], [-100, -110, -100, -110, -160]],
    [[20, 20, -20, -10, -20], [20, 20, -50, -30, -50], [-10, -10, -20, -10, -20],
     [-50, -100, -50, -110, -50], [-10, -10, -30, -10, -100]],
    [[0, -20, -10, -20, 0], [-30, -50, -30, -60, -30], [0, -20, -10, -20, 0],
     [-30, -90, -30, -110, -30], [-10, -20, -10, -20, -90]],
    [[-10, -10, -20, -10, -20], [-30, -30, -50, -30, -50], [-10, -10, -20, -10, -20],
     [-30, -120, -30, -110, -30], [-10, -10, -30, -10, -100]],
    [[0, -20, -10, -20, 0], [-30, -50, -30, -50, -30], [0, -20, -10, -20, 0],
     [-30, -150, -30, -150, -30], [-10, -20, -10, -20, -80]],
    [[20, 20, -10, -10, 0], [20, 20, -30, -30, -30], [0, -10, -10, -10, 0],
     [-30, -90, -30, -110, -30], [-10, -10, -10, -10, -80]],
]


def mm_interior_block(base, ag, ga, gg, uu):
    b = [[base] * 5 for _ in range(5)]
    b[1][3] = base + ag
    b[3][1] = base + ga
    b[3][3] = base + gg
    b[4][4] = base + uu
    return b


mismatch_interior = [mm_interior_block(0, -80, -100, -100, -60) for _ in range(2)] + \
                    [mm_interior_block(70, -80, -100, -100, -60) for _ in range(5)]
mismatch_interior_1n = [[[0] * 5 for _ in range(5)] for _ in range(2)] + \
                       [[[70] * 5 for _ in range(5)] for _ in range(5)]
mismatch_interior_23 = [
    mm_interior_block(0, -50, -110, -70, -30),
    mm_interior_block(0, 0, -120, -70, -30),
    mm_interior_block(70, 0, -110, -70, -30),
    mm_interior_block(70, -50, -110, -70, -30),
    mm_interior_block(70, 0, -110, -70, -30),
    mm_interior_block(70, -50, -110, -70, -30),
    mm_interior_block(70, 0, -110, -70, -30),
]

mismatch_ext = [
    [[-50, -110, -50, -140, -70], [-110, -110, -110, -160, -110], [-70, -150, -70, -150, -100],
     [-110, -130, -110, -140, -110], [-50, -150, -50, -150, -70]],
    [[-80, -140, -80, -140, -100], [-100, -150, -100, -140, -100], [-110, -150, -110, -150, -140],
     [-100, -140, -100, -160, -100], [-80, -150, -80, -150, -120]],
    [[-50, -80, -50, -50, -50], [-50, -100, -70, -50, -70], [-60, -80, -60, -80, -60],
     [-70, -110, -70, -80, -70], [-50, -80, -50, -80, -50]],
    [[-30, -30, -60, -60, -60], [-30, -30, -60, -60, -60], [-70, -100, -70, -100, -80],
     [-60, -80, -60, -80, -60], [-60, -80, -60, -80, -60]],
    [[-50, -80, -50, -80, -50], [-70, -100, -70, -110, -70], [-60, -80, -60, -80, -60],
     [-70, -110, -70, -120, -70], [-50, -80, -50, -80, -50]],
    [[-60, -80, -60, -80, -60], [-60, -80, -60, -80, -60], [-70, -100, -70, -100, -80],
     [-60, -80, -60, -80, -60], [-70, -100, -70, -100, -80]],
    [[-30, -30, -50, -50, -50], [-30, -30, -60, -50, -60], [-60, -80, -60, -80, -60],
     [-60, -80, -60, -80, -60], [-50, -80, -50, -80, -50]],
]
mismatch_multi = mismatch_ext

dangle5 = [
    [-10, -50, -30, -20, -10],
    [0, -20, -30, 0, 0],
    [-20, -30, -30, -40, -20],
    [-10, -30, -10, -20, -20],
    [-20, -30, -30, -40, -20],
    [-10, -30, -10, -20, -20],
    [0, -20, -10, 0, 0],
]
dangle3 = [
    [-40, -110, -40, -130, -60],
    [-80, -170, -80, -170, -120],
    [-10, -70, -10, -70, -10],
    [-50, -80, -50, -80, -60],
    [-10, -70, -10, -70, -10],
    [-50, -80, -50, -80, -60],
    [-10, -70, -10, -70, -10],
]

hairpin = [INF, INF, INF, 540, 560, 570, 540, 600, 550, 640, 650, 660, 670, 680, 690, 690, 700,
           710, 710, 720, 720, 730, 730, 740, 740, 750, 750, 750, 760, 760, 770]
bulge = [INF, 380, 280, 320, 360, 400, 440, 459, 470, 480, 490, 500, 510, 520, 530, 540, 540,
         550, 550, 560, 570, 570, 580, 580, 580, 590, 590, 600, 600, 600, 610]
interior = [INF, INF, 100, 100, 110, 200, 200, 210, 230, 240, 250, 260, 270, 280, 290, 290, 300,
            310, 310, 320, 330, 330, 340, 340, 350, 350, 350, 360, 360, 370, 370]

tetraloops = [("CAACGG", 550, 690), ("CCAAGG", 330, -1030), ("CCACGG", 370, -330),
              ("CCCAGG", 340, -890), ("CCGAGG", 350, -660), ("CCGCGG", 360, -750),
              ("CCUAGG", 370, -350), ("CCUCGG", 250, -1390), ("CUAAGG", 360, -760),
              ("CUACGG", 280, -1070), ("CUCAGG", 370, -660), ("CUCCGG", 270, -1290),
              ("CUGCGG", 280, -1070), ("CUUAGG", 350, -620), ("CUUCGG", 370, -1530),
              ("CUUUGG", 370, -680)]
triloops = [("CAACG", 680, 2370), ("GUUAC", 690, 1080)]
hexaloops = [("ACAGUACU", 280, -1680), ("ACAGUGAU", 360, -1140), ("ACAGUGCU", 290, -1280),
             ("ACAGUGUU", 180, -1540)]


def au(t):  # AU/GU closure (types index 2..6 => GU,UG,AU,UA,NS)
    return 1 if t >= 2 else 0


def b11(a, b):
    """1x1 mismatch bonus (a = 5' side nt, b = 3' side nt; 0=N,1=A,2=C,3=G,4=U)."""
    if a == 0 or b == 0:
        return 40
    if a == 3 and b == 3:
        return -190
    if a == 4 and b == 4:
        return -10
    if a == 1 and b == 1:
        return 40
    return 0


def mmbonus(a, b):
    if a == 0 or b == 0:
        return 0
    if (a, b) == (1, 3):
        return -80
    if (a, b) == (3, 1):
        return -100
    if (a, b) == (3, 3):
        return -100
    if (a, b) == (4, 4):
        return -60
    return 0


def gen_int11():
    out = []
    for t1 in range(7):
        for t2 in range(7):
            blk = [[50 + 70 * (au(t1) + au(t2)) + b11(a, b) for b in range(5)] for a in range(5)]
            out.append(blk)
    return out


def gen_int21():
    out = []
    for t1 in range(7):
        for t2 in range(7):
            for a in range(5):
                blk = [[230 + 70 * (au(t1) + au(t2)) + (mmbonus(a, c) // 2) + (-120 if (b == 3 and c == 3) else 0)
                        for c in range(5)] for b in range(5)]
                out.append(blk)
    return out


def gen_int22():
    out = []
    for t1 in range(6):
        for t2 in range(6):
            for a in range(1, 5):
                for b in range(1, 5):
                    blk = [[110 + 70 * (au(t1) + au(t2)) + mmbonus(a, d) + mmbonus(c, b)
                            for d in range(1, 5)] for c in range(1, 5)]
                    out.append(blk)
    return out


def fmt(v):
    return "   INF" if v >= INF else "%6d" % v


def scale_dH(v):
    if v >= INF:
        return v
    return int(v * 3)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    out_path = os.path.join(here, "..", "scanfold_b200", "params", "rna_turner2004_besteffort.par")
    L = []
    w = L.append
    w("## RNAfold parameter file v2.0")
    w("")
    w("/* scanfold-b200 BEST-EFFORT stand-in for rna_turner2004.par -- NOT the ViennaRNA file. */")
    w("/* besteffort=1 : recalled blocks unverified, int11/int21/int22 synthesised by rule,    */")
    w("/* enthalpies are placeholders (T != 37 refused).  See tools/gen_besteffort_par.py.     */")
    w("")

    def mat(name, rows, dh_rows=None):
        w("# " + name)
        for r in rows:
            w(" ".join(fmt(x) for x in r))
        w("")
        w("# " + name + "_enthalpies")
        for r in (dh_rows if dh_rows is not None else [[scale_dH(x) for x in r] for r in rows]):
            w(" ".join(fmt(x) for x in r))
        w("")

    def blocks(name, blks):
        for suffix, f in (("", lambda x: x), ("_enthalpies", scale_dH)):
            w("# " + name + suffix)
            for k, blk in enumerate(blks):
                w("/* block %d */" % k)
                for r in blk:
                    w(" ".join(fmt(f(x)) for x in r))
            w("")

    mat("stack", stack, stack_dH)
    blocks("mismatch_hairpin", mismatch_hairpin)
    blocks("mismatch_interior", mismatch_interior)
    blocks("mismatch_interior_1n", mismatch_interior_1n)
    blocks("mismatch_interior_23", mismatch_interior_23)
    blocks("mismatch_multi", mismatch_multi)
    blocks("mismatch_exterior", mismatch_ext)
    mat("dangle5", dangle5)
    mat("dangle3", dangle3)
    blocks("int11", gen_int11())
    blocks("int21", gen_int21())
    blocks("int22", gen_int22())
    for name, arr in (("hairpin", hairpin), ("bulge", bulge), ("interior", interior)):
        for suffix, f in (("", lambda x: x), ("_enthalpies", scale_dH)):
            w("# " + name + suffix)
            for k in range(0, 31, 10):
                w(" ".join(fmt(f(x)) for x in arr[k:k + 10]))
            w("")
    w("# ML_params")
    w("/* cu cu_dH cc cc_dH ci ci_dH */")
    w("     0      0    930   3000    -90   -220")
    w("")
    w("# NINIO")
    w("/* m m_dH max */")
    w("    60    320    300")
    w("")
    w("# Misc")
    w("/* DuplexInit dH TerminalAU dH LXC LXC_dH */")
    w("   410    360     50    370 107.856000 0")
    w("")
    for name, lst in (("Hexaloops", hexaloops), ("Tetraloops", tetraloops), ("Triloops", triloops)):
        w("# " + name)
        for s, e, h in lst:
            w("%s %6d %6d" % (s, e, h))
        w("")
    w("#END")
    with open(out_path, "w") as f:
        f.write("\n".join(L) + "\n")
    print("wrote", os.path.normpath(out_path), len(L), "lines")


if __name__ == "__main__":
    main()
