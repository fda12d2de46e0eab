#!/bin/bash
# partition-function kernels: parity, then windows/s by window length
TAG=${1:-r02n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_engine.py tests/test_gpu_configs.py -m gpu -x -q -k "partition or c3 or c5 or scan_" > $OUT/pytest_pf.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_pf.log
for W in 120 200 300 600; do SFB_PF_ENGINE=1 timeout 300 python tools/time_pf.py $W $((12000*120/W/ (W>200?4:1) )) 2>&1 | grep "rep 3"; done
timeout 300 python tools/time_pf.py 120 12000 2>&1 | grep "rep 3"
