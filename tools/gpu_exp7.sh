#!/bin/bash
OUT=gpurun_out/exp7; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
timeout 900 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; tail -3 $OUT/configs.err; cat $OUT/configs.jsonl | cut -c1-420
