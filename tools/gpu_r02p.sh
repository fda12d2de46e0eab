#!/bin/bash
# full GPU suite after the temperature / shim / flag work
TAG=${1:-r02p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 $OUT/pytest_gpu.log
