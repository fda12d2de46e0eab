#!/bin/bash
# ncu --set full capture of one kernel.  usage: bash tools/gpu_ncu.sh <tag> <kernel regex> [windows]
TAG=$1; K=$2; NWIN=${3:-2000}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $OUT/prof_$K \
    python bench.py --windows $NWIN --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_$K.log 2>&1
ls -la $OUT
