#!/bin/bash
# round-2 visit B: the blocked int32 kernel (mfe4.cu): parity, then folds/s by window length
TAG=${1:-r02b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "blocked or fold_kernels_agree or mfe_energy" > $OUT/pytest_mfe4.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_mfe4.log
for W in 301 450 600 1000; do timeout 300 python tools/time_mfe.py $W 40 2 2>&1 | tail -1; done
