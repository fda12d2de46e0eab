#!/bin/bash
OUT=gpurun_out/exp9; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "fold_kernels_agree or mfe_energy or scan" > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
for nw in 8 12 16; do SFB_MFE3_WARPS=$nw timeout 120 python tools/time_mfe.py 120 6000 3 2>&1 | tail -1 | sed "s/^/nw=$nw /"; done | tee $OUT/warps.log
timeout 300 python tools/time_mfe.py 200 600 2 2>&1 | tail -1 | tee $OUT/w200.log
timeout 300 python tools/time_mfe.py 40 6000 2 2>&1 | tail -1 | tee $OUT/w40.log
