#!/bin/bash
# quick correctness + timing of the fold kernels.  usage: bash tools/gpu_quick.sh
timeout 800 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "fold_kernels_agree or mfe_energy or flag_only" 2>&1 | tail -2
for nw in 8 12 16; do SFB_MFE3_WARPS=$nw timeout 120 python tools/time_mfe.py 120 6000 3 2>&1 | tail -1 | sed "s/^/nw=$nw /"; done
timeout 300 python tools/time_mfe.py 200 600 2 2>&1 | tail -1
timeout 300 python tools/time_mfe.py 40 6000 2 2>&1 | tail -1
