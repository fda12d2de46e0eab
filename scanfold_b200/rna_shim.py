"""Level-1 drop-in for the ViennaRNA `RNA` module -- exactly the surface the reference touches
(SURVEY.md 8b): RNA.md, RNA.fold_compound(seq[, md]) with .mfe() .pf() .centroid() .mean_bp_distance()
.hc_add_from_db() .sc_add_SHAPE_deigan(), plus RNA.fold / RNA.pf_fold used by the motif step.

Call sites replaced: ScanFold.py:212-215 (md), :494-544 (window folds), :725-731 (final window),
:1518-1539 (global refold), :1733-1741 (motifs); ScanFoldFunctions.py:774-789 (rna_folder).

Every fold goes to a backend object; the default backend is the CUDA engine (scanfold_b200.engine, one
fold per call -- correct but slow, the batched path is scanfold_b200.scan).  Tests install the CPU
oracle as backend to run the unmodified reference end to end and produce golden files.
`import RNA` resolves to this module through scanfold_b200/compat/RNA.py (put that directory on PYTHONPATH).
"""
import numpy as np

__version__ = "scanfold_b200-shim"

_backend = None


class EngineBackend:
    """folds on the GPU through the C-ABI (scanfold_b200.engine)"""

    def mfe(self, seq, hc, sc_stack, max_span, temperature=37.0):
        from . import engine
        e, pt = engine.fold_batch([seq], hc=[hc] if hc else None,
                                  sc=np.asarray(sc_stack, dtype=np.int32)[None, :] if sc_stack is not None else None,
                                  structure=True, max_span=max_span, temperature=temperature)
        return int(e[0]), engine.pair_table_to_dotbracket(pt[0])

    def pf(self, seq, hc, sc_stack, max_span, temperature):
        from . import engine
        res = engine.pf_batch([seq], hc=[hc] if hc else None,
                              sc=np.asarray(sc_stack, dtype=np.int32)[None, :] if sc_stack is not None else None,
                              want_bpp=True, temperature=temperature, max_span=max_span)
        return {"dG": float(res["dG"][0]), "ed": float(res["ed"][0]),
                "centroid": engine.pair_table_to_dotbracket(res["centroid"][0]), "bpp": res["bpp"][0]}

    def deigan(self, react1, m, b):
        from . import engine
        return engine.deigan(react1, m, b)


def set_backend(backend):
    global _backend
    _backend = backend


def backend():
    global _backend
    if _backend is None:
        _backend = EngineBackend()
    return _backend


class md:
    """RNA.md(): model details.  The reference sets only temperature (ScanFold.py:213, ScanFoldFunctions.py:777)
    and optionally max_bp_span (ScanFold.py:215); everything else stays at ViennaRNA's defaults."""

    def __init__(self):
        self.temperature = 37.0
        self.max_bp_span = -1
        self.dangles = 2
        self.noLP = 0
        self.noGU = 0


class fold_compound:
    def __init__(self, sequence, model=None, *unused):
        self.sequence = str(sequence).upper()
        self.length = len(self.sequence)
        self._temperature = float(model.temperature) if model is not None else 37.0
        span = int(model.max_bp_span) if model is not None else -1
        self._span = span if span > 0 else 0
        self._hc = None
        self._sc = None
        self._pf = None

    # ---- constraints
    def hc_add_from_db(self, constraint, options=None):
        s = str(constraint).rstrip("\n")
        if len(s) < self.length:
            s = s + "." * (self.length - len(s))
        self._hc = s[:self.length]
        return 1

    def sc_add_SHAPE_deigan(self, reactivities, m, b, options=None):
        """vrna_sc_add_SHAPE_deigan reads reactivities[1..n] (1-based); entries past the end of the given
        list are treated as missing (the reference passes a 0-based slice, Appendix B Q7)."""
        n = self.length
        r1 = np.full(n + 1, -999.0)
        src = list(reactivities)
        for p in range(1, n + 1):
            if p < len(src):
                r1[p] = float(src[p])
        self._sc = np.asarray(backend().deigan(r1, float(m), float(b)), dtype=np.int32)
        return 1

    def sc_add_SHAPE_zarringhalam(self, *args):
        raise TypeError("sc_add_SHAPE_zarringhalam needs (reactivities, b, default_value, shape_conversion); "
                        "the reference calls it with one argument (ScanFold.py:536), which fails in ViennaRNA too")

    # ---- folds
    def mfe(self):
        if abs(self._temperature - 37.0) > 1e-9:
            e, s = backend().mfe(self.sequence, self._hc, self._sc, self._span, temperature=self._temperature)
        else:
            e, s = backend().mfe(self.sequence, self._hc, self._sc, self._span)
        return [s, float(np.float32(e / 100.0))]

    def pf(self):
        self._pf = backend().pf(self.sequence, self._hc, self._sc, self._span, self._temperature)
        return ["", float(np.float32(self._pf["dG"]))]

    def _need_pf(self):
        if self._pf is None:
            raise RuntimeError("pf() must be called before centroid() / mean_bp_distance()")
        return self._pf

    def centroid(self):
        p = self._need_pf()
        dist = 0.0
        bpp = p.get("bpp")
        if bpp is not None:
            iu = np.triu_indices(self.length, 1)
            v = np.asarray(bpp)[iu]
            dist = float(np.where(v > 0.5, 1.0 - v, v).sum())
        return [p["centroid"], dist]

    def mean_bp_distance(self):
        return self._need_pf()["ed"]


def fold(sequence):
    """RNA.fold(seq) -> (structure, mfe) at the default model"""
    return fold_compound(sequence).mfe()


def pf_fold(sequence):
    """RNA.pf_fold(seq) -> (structure, ensemble energy); the reference discards the result (ScanFold.py:1736)"""
    fc = fold_compound(sequence)
    e = fc.pf()[1]
    return [fc.centroid()[0], e]


def duplexfold(*args):
    raise NotImplementedError("RNA.duplexfold (the experimental --lri scan, ScanFold.py:769-1034) is out of scope")


def PS_rna_plot_a(sequence, structure, filename, pre="", post=""):
    """ViennaRNA's PostScript layout engine is out of scope; write a placeholder so the
    motif step of the reference completes."""
    with open(filename, "w") as f:
        f.write("%%!PS-Adobe-3.0 EPSF-3.0\n%% scanfold_b200: structure layout not implemented\n%% %s\n%% %s\n"
                % (sequence, structure))
    return 1
