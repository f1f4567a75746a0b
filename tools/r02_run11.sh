#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run11.log; : > $out
run() { echo "=== $*" >> $out; env "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run BOBE_X=1 timeout 600 python tools/factor_ab.py check
run BOBE_X=2 timeout 600 python tools/factor_ab.py check
run BOBE_X=1 timeout 600 python tools/factor_ab.py time
run BOBE_GREEN_SMS=0 timeout 600 python tools/factor_ab.py time
run BOBE_MLL_MIN_PER_STREAM=8 timeout 600 python tools/factor_ab.py time
run BOBE_MLL_MIN_PER_STREAM=8 BOBE_GREEN_SMS=24 timeout 600 python tools/factor_ab.py time
run BOBE_MLL_MIN_PER_STREAM=8 BOBE_GREEN_SMS=8 timeout 600 python tools/factor_ab.py time
BOBE_MLL_MIN_PER_STREAM=8 python tools/timeline.py mll 8 > gpurun_out/r02_tl11_mll8_s1.txt 2>&1
grep -v "^n=" $out | grep "===\|factorize n=\|R=8\|R=1:\|R=16\|R=64\|check ok\|FAILED\|rror"
