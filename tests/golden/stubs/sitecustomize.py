"""GOLDEN-GENERATION STUB: make the reference deterministic without touching its source -- the two
ProcessPoolExecutor(12) pools per window (ScanFoldFunctions.py:140-144) become an in-process map, and the
unseeded `random` module is seeded."""
import concurrent.futures
import os
import random


class _SerialPool:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def map(self, fn, *iterables):
        return [fn(*args) for args in zip(*iterables)]


if os.environ.get("SCANFOLD_GOLDEN_SEED"):
    concurrent.futures.ProcessPoolExecutor = _SerialPool
    random.seed(int(os.environ["SCANFOLD_GOLDEN_SEED"]))
