#!/bin/bash
# round-2 visit K: full GPU suite (new flag goldens, whole-sequence fold), whole-sequence fold timing
TAG=${1:-r02k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -16 $OUT/pytest_gpu.log
timeout 900 python tools/time_refold.py 2000 10000 29903 2>&1 | tee $OUT/refold_times.txt
for W in 450 600; do timeout 300 python tools/time_mfe.py $W 40 2 2>&1 | tail -1; done
