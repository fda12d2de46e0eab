#!/bin/bash
# One GPU-box visit: parity tests, smoke, a quick bench, the ncu launch list and one full capture of the MFE kernel.
# usage (under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py --windows 2000 --steps 1 --warmup 1 --cpu-windows 16 > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "quick rc=$?"
cat $OUT/bench_quick.json | head -c 3000
timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json | head -c 3000
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --windows 1000 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mfe_fold -s 1 -c 1 -f -o $OUT/prof_mfe \
    python bench.py --windows 600 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
