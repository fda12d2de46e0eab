#!/bin/bash
OUT=gpurun_out/sanitize; mkdir -p $OUT
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 40 python tools/sanitize_small.py 40 120 > $OUT/racecheck.log 2>&1; tail -4 $OUT/racecheck.log; grep -o "[a-z0-9_]*\.cu:[0-9]*" $OUT/racecheck.log | sort | uniq -c
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "fold_kernels_agree" 2>&1 | tail -2
timeout 120 python tools/time_mfe.py 120 6000 3 2>&1 | tail -1
