#!/bin/bash
mkdir -p gpurun_out
BOBE_MLL_STREAMS=1 BOBE_LOOKAHEAD_MAX=0 BOBE_MLL_OWN_STREAM=0 BOBE_PDL=0 python tools/timeline.py mll 64 > gpurun_out/r02_tl7_mll64_serial.txt 2>&1
out=gpurun_out/r02_run7.log; : > $out
run() { echo "=== $*" >> $out; env "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run BOBE_LOOKAHEAD_MAX=0 timeout 600 python tools/factor_ab.py time
run BOBE_LOOKAHEAD_MAX=0 BOBE_FACTOR_PW=8 timeout 600 python tools/factor_ab.py time
run BOBE_LOOKAHEAD_MAX=0 BOBE_MLL_STREAMS=2 timeout 600 python tools/factor_ab.py time
run BOBE_LOOKAHEAD_MAX=0 BOBE_MLL_STREAMS=8 timeout 600 python tools/factor_ab.py time
run BOBE_TINY_MAX_TILES=0 timeout 600 python tools/factor_ab.py time
grep "===\|R=64\|R=16" $out
