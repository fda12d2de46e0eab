"""One FASTA record through scan -> ScanFold-Fold -> output files: the body of the record loop
ScanFold.py:266-1554 (without the opt-in global refold / LRI branches), batched.

  scan        scan.scan_record            (CUDA: folds, shuffles, partition function)
  accumulate  engine.Accumulator          (CUDA: per-nucleotide partner sums)
  fold step   foldstep.aggregate/compete  (host: argmin, competition)
  outputs     writers.*                   (host: byte-identical text files)
"""
import os

import numpy as np

from . import foldstep, writers


def fold_inputs(table):
    """per-window integers the accumulator takes: round(z*100), MFE in dcal (== round(MFE, 2) * 100), round(ED*100)"""
    z100 = np.rint(np.asarray(table.z) * 100.0).astype(np.int32)
    ed100 = np.rint(np.asarray(table.ed) * 100.0).astype(np.int32)
    return z100, np.asarray(table.mfe_dcal, dtype=np.int32), ed100


def partner_table_gpu(L, table):
    """ScanFold.py:564-677 + :1051-1139 on the device -> foldstep.PartnerTable"""
    from . import engine
    z100, mfe100, ed100 = fold_inputs(table)
    acc = engine.Accumulator(L, table.W, table.step, table.first_window, table.pair_tbl, z100, mfe100, ed100)
    try:
        return foldstep.table_from_compact(*acc.compact())
    finally:
        acc.close()


class RunNames:
    """file names of a record's outputs (ScanFold.py:370-388,1484-1500)"""

    def __init__(self, read_name, record_name, W, step, r, shuffle_type, name="UserInput", out6="./IGV_BP_Track",
                 final_partners_wig="./IGV_BP_Zavg_metrics", dbn1="Zavg_NoFilter", dbn2="Zavg_-1_pairs",
                 dbn3="Zavg_-2_pairs", dbn4="AllDBN.txt"):
        self.read_name, self.record_name, self.name = read_name, record_name, name
        self.outname = "%s.win_%d.stp_%d.rnd_%d.shfl_%s" % (read_name, W, step, r, shuffle_type)
        self.out6, self.final_partners_wig = out6, final_partners_wig
        self.dbn1, self.dbn2, self.dbn3, self.dbn4 = dbn1, dbn2, dbn3, dbn4


def write_scan_outputs(seq, table, names, temperature, step):
    """.out table and the four scan wig tracks (ScanFold.py:416,685,1495-1499); returns minz (:766)"""
    o = names.outname
    writers.write_out(o + ".out", names.read_name, seq, table, temperature)
    fin = table.final
    extra = (lambda key: [fin[key]]) if fin is not None else (lambda key: [])
    writers.write_wig(o + ".scan-MFE.wig", table.mfe.tolist() + extra("mfe"), step, names.name)
    z_all = table.z.tolist() + extra("z")
    writers.write_wig(o + ".scan-zscores.wig", z_all, step, names.name)
    writers.write_wig(o + ".scan-pvalue.wig", table.p.tolist() + extra("p"), step, names.name)
    writers.write_wig(o + ".scan-ED.wig", table.ed.tolist() + extra("ed"), step, names.name)
    return min(z_all)


def write_fold_outputs(seq, ptable, names, minz, step):
    """ScanFold-Fold: logs, CT / dbn / bp / wig / fasta files (ScanFold.py:1036-1500,1554).  Returns (agg, final)."""
    o = names.outname
    with open(o + ".ScanFold.log", "w") as log_total, open(o + ".ntPairCounts.log", "w") as sirna:
        sirna.write("i\tnuc\twindows\tbps\n")
        agg = foldstep.aggregate(ptable, seq, log_total, sirna)
    with open(o + ".ScanFold.FinalPartners.txt", "w") as log_win:
        fin = foldstep.compete(agg, seq, log_win)
    n = agg.n_nt
    covered = seq[:n]
    dbn_text = []
    for path, filt, title in ((names.dbn1, 10.0, "NoFilter"), (names.dbn2, -1.0, "Zavg_-1"), (names.dbn3, -2.0, "Zavg_-2")):
        partner = writers.write_ct(path + ".ct", fin, seq, filt, names.name)
        writers.write_dbn(path + ".dbn", title, covered, partner)
        dbn_text.append(open(path + ".dbn").read())
    writers.write_bp(names.out6 + "." + o + ".bp", fin.i, fin.j, fin.z, names.name, minz)
    writers.write_wig_dict(names.final_partners_wig + "." + o + ".wig", fin.z, names.name, step)
    key = np.arange(1, n + 1)
    writers.write_bp(names.out6 + "." + o + "." + names.record_name + ".ALL.bp", key, agg.part, agg.mean_z, names.name,
                     minz)
    writers.write_fasta(names.name + "." + o + ".fa", seq, names.name)
    with open(names.dbn4, "w") as f:      # `cat` of the three dbn files (ScanFold.py:1554)
        f.write("".join(dbn_text))
    return agg, fin
