"""`import RNA` -> the scanfold_b200 Level-1 shim (put this directory on PYTHONPATH to run the unmodified
reference ScanFold.py on the CUDA engine)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from scanfold_b200.rna_shim import *  # noqa: F401,F403,E402
from scanfold_b200.rna_shim import __version__  # noqa: F401,E402
