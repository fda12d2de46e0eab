#!/usr/bin/env python
"""Drop-in entry point: same command line and output table as the reference's legacy ScanFold-Scan.py, folds on the B200 engine."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from scanfold_b200.legacy_scan import main  # noqa: E402

if __name__ == "__main__":
    main()
