#!/bin/bash
OUT=gpurun_out/exp8; mkdir -p $OUT
for lib in scanfold_b200/libscanfold_b200.so scanfold_b200/libsfb_fake.so; do for nw in 8 12; do SFB_DEBUG_NO_REDO=1 SFB_LIB=/root/repo/$lib SFB_MFE3_WARPS=$nw timeout 120 python tools/time_mfe.py 120 6000 3 2>&1 | tail -1 | sed "s#^#$lib nw=$nw #"; done; done | tee $OUT/variants.log
