#!/bin/bash
# round-2 closing evidence: all GPU tests, smoke, bench (ours + reference arm), launch list, per-config throughput,
# ncu --set full of mfe3_kernel and pf2_kernel, fold rates by window length
TAG=${1:-r02y}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 1200 $OUT/bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; head -c 400 $OUT/bench_reference.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --windows 3000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launches.log 2>&1
for k in mfe3_kernel pf2_kernel; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/prof_$k \
    python bench.py --windows 2000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_$k.log 2>&1
done
for w in 64 120 200 300 301 600; do timeout 300 python tools/time_mfe.py $w $([ $w -gt 300 ] && echo 400 || echo 3000) 3 2>&1 | tail -1; done > $OUT/fold_rates.txt; cat $OUT/fold_rates.txt
for w in 120 200 300; do timeout 300 python tools/time_pf.py $w $([ $w -gt 120 ] && echo 2000 || echo 12000) 2>&1 | grep "rep 3"; done > $OUT/pf_rates.txt; cat $OUT/pf_rates.txt
timeout 1500 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; cut -c1-330 $OUT/configs.jsonl
ls -la $OUT
