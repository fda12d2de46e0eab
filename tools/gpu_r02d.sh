#!/bin/bash
# ncu --set full of one mfe4_block_kernel launch (a middle block diagonal) at W=600
TAG=${1:-r02d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mfe4_block_kernel -s 8 -c 1 -f -o $OUT/prof_mfe4_block \
    python tools/time_mfe.py 600 40 1 > $OUT/ncu_mfe4.log 2>&1; tail -2 $OUT/ncu_mfe4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_w600.csv \
    python tools/time_mfe.py 600 40 1 > $OUT/ncu_launches.log 2>&1
ls -la $OUT
