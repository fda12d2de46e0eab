#!/bin/bash
OUT=gpurun_out/exp12; mkdir -p $OUT
timeout 800 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "fold_kernels_agree or mfe_energy or flag_only" > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
timeout 300 python tools/time_mfe.py 300 600 2 2>&1 | tail -1 | tee $OUT/w300.log
timeout 300 python tools/time_mfe.py 250 600 2 2>&1 | tail -1 | tee $OUT/w250.log
timeout 200 python tools/time_mfe.py 120 6000 2 2>&1 | tail -1 | tee $OUT/w120.log
