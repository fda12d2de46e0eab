"""The C-ABI shared library builds, loads without a GPU and exports every symbol include/scanfold_b200.h declares.
No compute call is made here; on a box without a device sfb_init must fail loudly (no CPU fallback)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "scanfold_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from scanfold_b200 import build, engine
    build.build()
    lib = engine.load_library()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(engine.EXPORTS) == syms
    assert lib.sfb_version() == 100


def test_no_cpu_fallback_without_device():
    import torch
    from scanfold_b200 import engine
    if torch.cuda.is_available():
        return
    lib = engine.load_library()
    rc = lib.sfb_init(0, None)
    assert rc != 0
    assert b"no CPU fallback" in lib.sfb_last_error() or b"CUDA" in lib.sfb_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scanfold_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                text = open(os.path.join(dp, f)).read()
                assert "sf_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
