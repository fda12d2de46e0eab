#!/bin/bash
# round-2 closing check of the committed state: all GPU tests, smoke, bench (ours + reference arm), launch list, CLI wall time
TAG=${1:-r02z}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 600 $OUT/bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; head -c 300 $OUT/bench_reference.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --windows 3000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launches.log 2>&1
timeout 300 python tools/time_cli.py > $OUT/cli_c2_walltime.txt 2>&1; grep "CLI wall" $OUT/cli_c2_walltime.txt
ls -la $OUT
