#!/usr/bin/env python
"""Draws the distribution fixture for the device dinucleotide shuffle from the REFERENCE's own dinuclShuffle
(/root/reference/ScanFoldFunctions.py:255-277, imported unmodified with the RNA / Bio stubs of this directory).

For a few short sequences whose dinucleotide-shuffle space is small, 200,000 shuffles each are drawn with a seeded
`random`, and the number of times every distinct outcome appeared is stored in dishuffle_counts.json.
tests/test_gpu_configs.py draws the same number of shuffles on the device (Philox, Altschul-Erikson in shuffle.cu)
and compares the two frequency tables with a chi-square test.

Run here (needs /root/reference):  python tests/golden/make_shuffle_fixture.py
"""
import collections
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "stubs"))
sys.path.insert(0, "/root/reference")

SEQS = ["ACGUACGUUGCAAC", "GGAUCCGAUUAGCC", "AUAUGCGCAUAUGC", "ACGUUGCAGUCAGUACGU", "GAUCGAUCGGAUCUAG"]
N = 200000


def main():
    import ScanFoldFunctions as SFF          # the reference module, unmodified
    random.seed(20261017)
    out = []
    for s in SEQS:
        cnt = collections.Counter(SFF.dinuclShuffle(s) for _ in range(N))
        out.append({"seq": s, "draws": N, "counts": dict(sorted(cnt.items()))})
        print(s, "distinct outcomes:", len(cnt), "min count", min(cnt.values()))
    json.dump(out, open(os.path.join(HERE, "dishuffle_counts.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
