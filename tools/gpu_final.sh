#!/bin/bash
# round-end visit: parity tests, smoke, bench (+ reference arm), launch list, per-config throughput
TAG=${1:-r01x}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 2600 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --windows 3000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launches.log 2>&1
timeout 900 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; cut -c1-260 $OUT/configs.jsonl
