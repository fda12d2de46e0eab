"""Motif extraction and per-motif refold -- the step that follows the ScanFold-Fold outputs in a stock run
(ScanFold.py:1557-1781): every top-level helix of the Zavg < -2 structure is cut out, refolded with
its ScanFold pairs as hard constraints, and scored against 100 shuffles.

Per motif the reference makes the same calls as for one scan window (fold with hc, partition function with hc,
unconstrained native + 100 shuffled folds, z-score, p-value; ScanFold.py:1724-1756), so each motif is ONE
single-window scan of the CUDA engine (scan.scan_record with W = motif length).  Text outputs follow the
reference byte for byte: <name>_motif_<n>.dbn (:1763-1764), its .ct (dbn2ct, ScanFoldFunctions.py:918-1012) and one
gff3 line per motif (:1777-1778).  The PostScript plots (:1776) need ViennaRNA's layout engine and are not written.
"""
import numpy as np

from . import scan

SUB_RANDOMIZATIONS = 100          # sub_randomizations, ScanFold.py:1564 (the scramble call hard-codes 100, :1746)


class Motif:
    def __init__(self, number, sequence, structure, i, j):
        self.number, self.sequence, self.structure, self.i, self.j = number, sequence, structure, i, j


def extract_motifs(structure_line, seq, log=print):
    """ScanFold.py:1585-1712.  structure_line: line 3 of the Zavg_-2 dbn file INCLUDING its newline (the loops run to
    len-1).  Returns the motifs with 0-based inclusive coordinates (i, j) into the record.

    Quirks kept: the bond order of a nucleotide is the nesting depth AFTER it (so a helix opens where the depth
    becomes 1 on '(' and closes where it becomes 0 on ')'); '<' '>' '{' '}' never change the depth but open / close a
    motif when they sit at depth 1 / 0; an opener at index 0 is lost because the start coordinate is looked up in the
    1-based nucleotide dictionary (KeyError -> "EXCEPT", :1698-1701)."""
    structure = list(structure_line)
    n = len(structure) - 1
    depth = 0
    refold = {}                                   # index -> (bond_order, structure char)
    for m in range(max(n, 0)):
        ch = structure[m]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch not in ".<>{}":
            continue
        refold[m] = (depth, ch)
    starts, ends = [], []
    for j in range(max(n, 0)):
        if j not in refold:
            log(j, "EXCEPT")
            continue
        order, ch = refold[j]
        if order == 1 and ch in "(<{":
            if j == 0 or j > len(seq):            # nuc_dict holds coordinates 1..L only
                log(j, "EXCEPT")
                continue
            starts.append(j)                      # nuc_dict[j].coordinate == j
        elif order == 0 and ch in ")>}":
            ends.append(j + 1)                    # nuc_dict_refold[j].coordinate == j + 1
    motifs = []
    for l in range(len(starts)):
        s, e = starts[l], ends[l] - 1             # IndexError if the closers run out, as in the reference (:1703)
        frag = "".join(seq[k] for k in sorted(refold) if s <= k <= e)
        st = "".join(refold[k][1] for k in sorted(refold) if s <= k <= e)
        motifs.append(Motif(l + 1, frag, st, s, e))
    return motifs


def dbn_to_ct_text(sequence, structure):
    """dbn2ct, ScanFoldFunctions.py:918-1012: header `len-1 <tab> SequenceID`, one line per '.', '(' or ')'
    nucleotide whose running depth is >= 0 (pseudoknot characters get no line)."""
    n = len(sequence)
    if n != len(structure):
        raise ValueError("ERROR structure and sequence not same length.")
    pair = [0] * n
    stack = []
    for k, ch in enumerate(structure):
        if ch == "(":
            stack.append(k)
        elif ch == ")" and stack:
            a = stack.pop()
            pair[a], pair[k] = k + 1, a + 1
    out = ["%d\tSequenceID\n" % (n - 1)]
    depth = 0
    for k, ch in enumerate(structure):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch not in ".<>{}":
            continue
        if depth >= 0 and ch in ".()":
            out.append("%d %s %d %d %d %d\n" % (k + 1, sequence[k], k, k + 2, pair[k] if ch != "." else 0, k + 1))
    return "".join(out)


def refold_motif(m, shuffle_type, temperature, seed, parity_shuffles=None):
    """One motif = one single-window scan: native fold and partition function under the motif's hard constraints,
    unconstrained native + 100 shuffled folds, z-score / p-value (ScanFold.py:1724-1756)."""
    frag = m.sequence.upper().replace("T", "U")
    W = len(frag)
    par = None
    if parity_shuffles is not None:
        par = np.ascontiguousarray(parity_shuffles, dtype=np.uint8).reshape(1, SUB_RANDOMIZATIONS, W)
    # the constrained fold and its partition function use a default compound (37 C, ScanFold.py:1733-1741); only the
    # background energies() call gets the -t temperature (:1748)
    t = scan.scan_record(frag, W, 1, SUB_RANDOMIZATIONS, shuffle_type=shuffle_type, seed=seed, parity_shuffles=par,
                         temperature=37.0, background_temperature=temperature, hc=m.structure, final_window=False)
    return {"structure": t.structure(0), "mfe": float(t.mfe[0]), "z": float(t.z[0]), "p": float(t.p[0]),
            "ed": float(t.ed[0])}


def write_motif_outputs(motifs, results, name, gff_path):
    """<name>_motif_<n>.dbn / .ct and the gff3 table (ScanFold.py:1722,1763-1778)"""
    with open(gff_path, "w") as se:
        for m, r in zip(motifs, results):
            base = "%s_motif_%d" % (name, m.number)
            with open(base + ".dbn", "w") as f:
                f.write(">%s_coordinates:%d-%d\n%s\n%s" % (base, m.i, m.j, m.sequence, r["structure"]))
            with open(base + ".ct", "w") as f:
                f.write(dbn_to_ct_text(m.sequence.strip(), r["structure"].strip()))
            attrs = "motif_%d;sequence=%s;structure=%s;refoldedMFE=%s;MFE(kcal/mol)=%s;z-score=%s;ED=%s" % (
                m.number, m.sequence, m.structure, r["structure"], str(r["mfe"]), str(r["z"]), str(r["ed"]))
            se.write("%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % (name, ".", "RNA_sequence_secondary_structure", str(m.i + 1),
                                                           str(m.j + 1), ".", ".", ".", attrs))


def run(seq, dbn2_path, name, gff_path, shuffle_type="mono", temperature=37.0, seed=42, parity=None, log=print):
    """The whole step for one record.  parity: mapping with `motif_shuffles_<n>` arrays (the shuffles a reference run
    drew), or None for device shuffles keyed by (seed, motif number)."""
    lines = open(dbn2_path).readlines()
    motifs = extract_motifs(lines[2], seq, log)
    results = []
    for m in motifs:
        par = None
        if parity is not None:
            key = "motif_shuffles_%d" % m.number
            if key not in parity:
                raise KeyError("parity file has no %s (motif %d-%d)" % (key, m.i, m.j))
            par = parity[key]
        results.append(refold_motif(m, shuffle_type, temperature, seed + 7919 * m.number, par))
    write_motif_outputs(motifs, results, name, gff_path)
    return motifs, results
