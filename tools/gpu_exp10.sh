#!/bin/bash
OUT=gpurun_out/exp10; mkdir -p $OUT
for lib in libsfb_base.so libscanfold_b200.so; do SFB_LIB=/root/repo/scanfold_b200/$lib timeout 200 python tools/time_pf.py 120 6000 2>&1 | tail -3 | sed "s/^/$lib /"; done | tee $OUT/pf.log
