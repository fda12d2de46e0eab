"""z-score and p-value of a window against its shuffled background -- batched restatement of
zscore_function / pvalue_function (ScanFoldFunctions.py:741-751, :727-738) as called at
ScanFold.py:553-562, for all windows at once.

Quirks kept on purpose:
  Q1  the background mean runs over shuffles 1..r-1 (energy_list[1:randomizations]) while the standard
      deviation is the SAMPLE stdev over native + all r shuffles; sd == 0 gives z = 0.0
  Q2  p = #{E < E_native} / (r + 1) with a strict '<'
  Q3  energies enter as C floats (ViennaRNA returns (float)e/100.), the arithmetic is Python's
      `statistics` on those doubles, and both results are round(x, 2)

The bulk is vectorised float64; any window whose z lands within 1e-7 of a rounding boundary is
recomputed with the `statistics` module itself, so the rounded value is always the reference's.
"""
import statistics

import numpy as np


def energy_to_float(e_dcal):
    """int dcal -> the double holding ViennaRNA's C float (float)e/100.  (vrna_mfe return value)"""
    return (np.asarray(e_dcal, dtype=np.float64) / 100.0).astype(np.float32).astype(np.float64)


def round_energy(e_dcal):
    """round(MFE, 2) of ScanFold.py:501 -- the nearest double to e/100"""
    return np.asarray(e_dcal, dtype=np.float64) / 100.0


def _zscore_exact(row, r):
    """the reference's own arithmetic on one energy list (ScanFoldFunctions.py:741-751)"""
    lst = [float(x) for x in row]
    sd = statistics.stdev(lst)
    if sd == 0:
        return 0.0
    return (lst[0] - statistics.mean(lst[1:r])) / sd


def zscore_pvalue(native_dcal, shuffle_dcal):
    """native_dcal [n] int32 (energy_list[0], the UNCONSTRAINED native refold, Q4), shuffle_dcal [n, r]
    -> (z [n] float64 rounded to 2 decimals, p [n] float64 rounded to 2 decimals)"""
    nat = np.asarray(native_dcal)
    sh = np.asarray(shuffle_dcal)
    n, r = sh.shape
    if r < 2:
        raise ValueError("the reference's z-score needs at least 2 randomizations "
                         "(statistics.mean of energy_list[1:r] is empty otherwise)")
    E = np.empty((n, r + 1), dtype=np.float64)
    E[:, 0] = energy_to_float(nat)
    E[:, 1:] = energy_to_float(sh)
    mean_sh = E[:, 1:r].sum(axis=1) / (r - 1)
    mean_all = E.sum(axis=1) / (r + 1)
    ss = ((E - mean_all[:, None]) ** 2).sum(axis=1)
    sd = np.sqrt(ss / r)
    # sd == 0 must be decided exactly: all r+1 energies identical
    const = (sh == nat[:, None]).all(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        zraw = np.where(const, 0.0, (E[:, 0] - mean_sh) / np.where(const, 1.0, sd))
    z100 = zraw * 100.0
    frac = np.abs(z100 - np.floor(z100) - 0.5)
    risky = np.nonzero((frac < 1e-7) & ~const)[0]
    z = np.rint(z100) / 100.0
    for k in risky:
        z[k] = round(_zscore_exact(E[k], r), 2)
    # np.rint keeps the sign of zero, like round(): round(-0.001, 2) is -0.0 and prints "-0.0"
    below = (E[:, 1:] < E[:, :1]).sum(axis=1)
    ptab = np.array([round(c / float(r + 1), 2) for c in range(r + 2)])
    return z, ptab[below]
