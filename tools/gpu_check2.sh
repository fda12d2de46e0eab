#!/bin/bash
# tests + quick bench + ncu capture of the warp-per-fold kernel.  usage: bash tools/gpu_check2.sh <tag> [full]
TAG=${1:-r01x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/pytest_gpu.log; cat $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --cpu-windows 16 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json | head -c 2500; tail -3 $OUT/bench.err
if [ "$2" == "full" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --windows 3000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mfe2_kernel -s 1 -c 1 -f -o $OUT/prof_mfe2 \
    python bench.py --windows 3000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full.log 2>&1
ls -la $OUT
fi
