#!/bin/bash
# numbers for the docs: shim test, bench (ours + reference arm), per-config throughput, launch list
TAG=${1:-r02q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_shim.py tests/test_viennarna_conditional.py -m gpu -q -rs > $OUT/pytest_shim.log 2>&1; tail -4 $OUT/pytest_shim.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 1200 $OUT/bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; head -c 400 $OUT/bench_reference.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --windows 3000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launches.log 2>&1
timeout 1500 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; cut -c1-330 $OUT/configs.jsonl
