"""GOLDEN-GENERATION STUB (test infrastructure): `import RNA` for the unmodified reference, backed by the CPU
oracle through scanfold_b200.rna_shim, recording every fold so make_golden.py can rebuild per-window arrays."""
import atexit
import json
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, _ROOT)
from scanfold_b200 import rna_shim  # noqa: E402
from scanfold_b200.rna_shim import *  # noqa: F401,F403,E402
from oracle import oracle as _O  # noqa: E402

_trace = []


class _OracleBackend:
    def mfe(self, seq, hc, sc_stack, max_span, temperature=37.0):
        e, s = _O.mfe(seq, hc=hc, sc_stack=sc_stack, max_span=max_span, temperature=temperature)
        _trace.append({"op": "mfe", "seq": seq, "hc": hc, "sc": None if sc_stack is None else [int(x) for x in sc_stack],
                       "e": int(e), "s": s})
        return e, s

    def pf(self, seq, hc, sc_stack, max_span, temperature):
        r = _O.pf(seq, hc=hc, sc_stack=sc_stack, max_span=max_span, temperature=temperature, want_bpp=True)
        _trace.append({"op": "pf", "seq": seq, "hc": hc, "sc": None if sc_stack is None else [int(x) for x in sc_stack],
                       "dG": r["dG"], "ed": r["ed"], "centroid": r["centroid"]})
        return r

    def deigan(self, react1, m, b):
        return _O.deigan(react1, m, b)


rna_shim.set_backend(_OracleBackend())


def _dump():
    path = os.environ.get("SCANFOLD_GOLDEN_TRACE")
    if path:
        with open(path, "w") as f:
            json.dump(_trace, f)


atexit.register(_dump)
