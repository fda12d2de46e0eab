#!/bin/bash
# One GPU-box visit: parity tests, smoke, full bench (+reference arm), launch list, ncu captures of the kernels named.
# usage (under gpurun): bash tools/gpu_check4.sh <tag> [kernel regexes...]
TAG=${1:-r01x}; shift; KERNELS=${@:-mfe3_kernel pf2_kernel}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
fi
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 2600 $OUT/bench.json
if [ -z "$SKIP_REF" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --windows 3000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launches.log 2>&1
for k in $KERNELS; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/prof_$k \
    python bench.py --windows 2000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_$k.log 2>&1
done
ls -la $OUT
