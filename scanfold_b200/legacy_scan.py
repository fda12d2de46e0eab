"""The legacy two-step surface, scan half: the command line, the arithmetic quirks and the `.txt` table of the reference's
ScanFold-Scan.py, with every fold of a record in ONE batched call into the CUDA engine.

What differs from the scan of ScanFold.py (and is kept, line numbers into /root/reference/ScanFold-Scan.py):
  * defaults -s 10 -r 50; one output file `<fasta>.forward.win_W.stp_S.rnd_R.shfl_TYPE.txt` for all records (:71), a
    header line per record (:335), ten tab-separated columns per window (:425); no final-window block (:341)
  * the native fold / partition function use the model with `-t` (:74-75,:364), the background -- the unconstrained native
    refold and every shuffle -- is folded by RNA.fold at the default 37 C (:244-246)
  * z = (E_native - mean(E[1:r])) / numpy.std(E) with the POPULATION standard deviation over native + r shuffles and the
    mean over the first r-1 shuffles (:232-242); a zero deviation prints "#DIV/0!"; p = #{E < E_native} / (r + 1) (:218-229)
  * a window that is literally 120 N is not folded: MFE 0, z "#DIV/0", ED 0, p 0, 120 dots (:355-361)
  * `-c` takes line 3 of the file INCLUDING its newline, one character per nucleotide (:300-317)
The arithmetic on the energy lists is numpy's own (np.mean / np.std on the doubles that hold ViennaRNA's C floats), so the
printed values are the reference's digit for digit.
"""
import argparse
import sys

import numpy as np

from . import engine, scan, stats
from .cli import read_fasta

ALL_N = "N" * 120


def build_parser():
    """the flags of ScanFold-Scan.py:41-60, same names and defaults"""
    p = argparse.ArgumentParser()
    p.add_argument("-i", "--filename", type=str, help="input filename")
    p.add_argument("-s", type=int, default=10, help="step size")
    p.add_argument("-w", type=int, default=120, help="window size")
    p.add_argument("-r", type=int, default=50, help="randomizations")
    p.add_argument("-t", type=int, default=37, help="Folding temperature")
    p.add_argument("-type", type=str, default="mono", help="randomization type")
    p.add_argument("-p", "--print_to_screen", action="store_true", help="print to screen option (default off)")
    p.add_argument("--print_random", type=str, default="off", help="print to screen option (default off)")
    p.add_argument("-c", "--constraints", type=str, help="optional | input constraint file")
    p.add_argument("--seed", type=int, default=42, help="[scanfold_b200] Philox seed of the device shuffles")
    p.add_argument("--parity_shuffles", type=str, default=None,
                   help="[scanfold_b200] .npz with `shuffles` [windows, r, W] uint8 to fold instead of device shuffles")
    p.add_argument("--params", type=str, default=None, help="[scanfold_b200] ViennaRNA parameter file")
    p.add_argument("--gpu", type=int, default=0, help="[scanfold_b200] CUDA device ordinal")
    return p


def output_name(fasta, W, step, r, stype):
    """ScanFold-Scan.py:71"""
    return "%s.forward.win_%s.stp_%s.rnd_%s.shfl_%s.txt" % (fasta, W, step, r, stype)


def window_starts(L, W, step):
    """`while i == 0 or i <= (length - window_size)` (:341) for a record of at least W nucleotides"""
    return list(range(0, L - W + 1, step))


def legacy_stats(native_dcal, shuffle_dcal, r):
    """zscore_function / pscore_function (:218-242) on one window's energy list -> (energy_list, z, p) as the reference
    holds them: z is round(numpy.float64, 2) or the string "#DIV/0!", p a Python float"""
    energy_list = [float(x) for x in stats.energy_to_float(np.concatenate([[native_dcal], shuffle_dcal]))]
    sd = np.std(energy_list)
    if sd != 0:
        z = round((energy_list[0] - np.mean(energy_list[1:r])) / sd, 2)
    else:
        z = "#DIV/0!"
    below = sum(1 for e in energy_list if float(e) < float(energy_list[0]))
    p = round(float(float(below) / float(len(energy_list))), 2)
    return energy_list, z, p


def format_record(read_name, seq, W, step, r, temperature, table, hc_line=None, print_to_screen=False,
                  print_random="off", out=None):
    """Header and rows of one record (:335,:425) from the engine's per-window arrays (`table`: scan.WindowTable of the
    windows at window_starts, without a final-window set).  Returns the text for the output file."""
    out = out or sys.stdout
    rows = ["i\tj\tTemperature\tNative_dG\tZ-score\tP-score\tEnsembleDiversity\tSequence\tStructure\tCentroid\t" + read_name + "\n"]
    structures = engine.pair_tables_to_dotbrackets(table.pair_tbl) if len(table) else []
    centroids = engine.pair_tables_to_dotbrackets(table.centroid_tbl) if len(table) else []
    for k, i in enumerate(window_starts(len(seq), W, step)):
        start, end = i + 1, i + W
        frag = seq[i:i + W].replace("T", "U").replace("t", "u")   # Seq.transcribe()
        if frag == ALL_N:
            mfe, z, ed, p = 0, "#DIV/0", 0, 0
            structure = centroid = "." * 120
        else:
            mfe, ed = float(table.mfe[k]), float(table.ed[k])
            structure, centroid = structures[k], centroids[k]
            energy_list, z, p = legacy_stats(table.native_unconstrained_dcal[k], table.shuffle_dcal[k], r)
            if print_random == "on":
                print(energy_list, file=out)
        if print_to_screen:
            head = "%s\t%s\t%s\t%s\t%s\t%s\t%s\n%s\n" % (start, end, temperature, mfe, z, p, ed, frag)
            if hc_line is not None:
                head += "".join(hc_line[start - 1:end]) + "\n"
            print(head + structure + "\n" + centroid + "\n", file=out)
        rows.append("%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % (start, end, temperature, mfe, z, p, ed, frag, structure, centroid))
    return "".join(rows)


def scan_legacy_record(seq, W, step, r, temperature, stype, hc_line=None, seed=42, parity_shuffles=None):
    """every window of one record on the engine: native fold + PF at `temperature` (under the constraint line if any),
    background at 37 C"""
    rna = seq.replace("T", "U").replace("t", "u")
    hc = None
    if hc_line is not None:
        hc = "".join(hc_line[:len(seq)])
    return scan.scan_record(rna, W=W, step=step, r=r, shuffle_type=stype, seed=seed, parity_shuffles=parity_shuffles,
                            temperature=float(temperature), hc=hc, final_window=False,
                            background_temperature=37.0 if int(temperature) != 37 else None)


def main(argv=None, out=None):
    out = out or sys.stdout
    args = build_parser().parse_args(argv)
    if args.type not in ("mono", "di"):
        print("Shuffle type not properly designated; please input \"di\" or \"mono\"", file=out)
    engine.init(args.gpu, args.params)
    if engine.params_besteffort():
        sys.stderr.write("scanfold_b200: WARNING: folding with the best-effort stand-in parameter file; pass --params / "
                         "SCANFOLD_PARAMS=rna_turner2004.par for ViennaRNA's energies\n")
    W, step, r = int(args.w), int(args.s), int(args.r)
    parity = np.load(args.parity_shuffles) if args.parity_shuffles else None
    with open(output_name(args.filename, W, step, r, args.type), "w") as w:
        hc_line = None
        for n_rec, (read_name, seq) in enumerate(read_fasta(args.filename)):
            print("Scanning sequence " + read_name + "\nSequence Length: " + str(len(seq)) + "nt long.", file=out)
            if args.constraints is not None:
                print("Considering constraint input", file=out)
                with open(args.constraints) as f:
                    hc_line = list(f.readlines()[2])
                print("Constraint list is " + str(len(hc_line) - 1) + "nt long.", file=out)
                if len(hc_line) - 1 != len(seq):
                    raise ValueError("Error detected. Sequence and Constraints must be same length.")
            if len(seq) < W:
                continue
            key = "shuffles" if n_rec == 0 else "shuffles_%d" % n_rec
            table = scan_legacy_record(seq, W, step, r, args.t, args.type, hc_line=hc_line, seed=args.seed,
                                       parity_shuffles=parity[key] if parity is not None else None)
            w.write(format_record(read_name, seq, W, step, r, int(args.t), table, hc_line=hc_line,
                                  print_to_screen=args.print_to_screen, print_random=str(args.print_random), out=out))


if __name__ == "__main__":
    main()
