#!/bin/bash
out=gpurun_out/r02_run2.log
mkdir -p gpurun_out; : > $out
run() { echo "=== $*" >> $out; env "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
tools/leaf_bench > gpurun_out/r02_leaf_bench3.txt 2>&1
run BOBE_X=1 timeout 600 python tools/factor_ab.py check
run BOBE_LOOKAHEAD_MAX=0 timeout 600 python tools/factor_ab.py check
run BOBE_X=1 timeout 600 python tools/factor_ab.py time
run BOBE_TINY_MAX_TILES=0 timeout 600 python tools/factor_ab.py time
run BOBE_MLL_MIN_PER_STREAM=8 timeout 600 python tools/factor_ab.py time
run BOBE_MLL_MIN_PER_STREAM=2 timeout 600 python tools/factor_ab.py time
python tools/timeline.py factor 2000 > gpurun_out/r02_tl2_factor.txt 2>&1
python tools/timeline.py mll 8 > gpurun_out/r02_tl2_mll8.txt 2>&1
grep -v "^n=" $out | tail -60
