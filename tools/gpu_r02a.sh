#!/bin/bash
# round-2 visit A: full GPU suite (new config-geometry and shuffle-statistics tests), racecheck, bench, ncu of the
# int32 long-window kernel (what does mfe_fold_kernel stall on at W=600?)
TAG=${1:-r02a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -25 $OUT/pytest_gpu.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 40 python tools/sanitize_small.py 40 120 > $OUT/racecheck.log 2>&1; tail -3 $OUT/racecheck.log; grep -o "[a-z0-9_]*\.cu:[0-9]*" $OUT/racecheck.log | sort | uniq -c
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfe_fold_kernel -c 1 -f -o $OUT/prof_mfe_w600 \
    python tools/time_mfe.py 600 8 1 > $OUT/ncu_mfe_w600.log 2>&1; tail -2 $OUT/ncu_mfe_w600.log
for W in 300 450 600; do timeout 300 python tools/time_mfe.py $W 24 2 2>&1 | tail -1; done
ls -la $OUT
