#!/bin/bash
# full bench + launch list + ncu captures of the three fold kernels.   usage: bash tools/gpu_check3.sh <tag>
TAG=${1:-r01x}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 2200 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --windows 3000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launches.log 2>&1
for k in mfe2_kernel pf_kernel mfe_fold_kernel; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/prof_$k \
    python bench.py --windows 2000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_$k.log 2>&1
done
ls -la $OUT
