#!/bin/bash
# A/B of the factorisation schemes (round 2).  Output: gpurun_out/r02_factor_sweep.log
out=gpurun_out/r02_factor_sweep.log
mkdir -p gpurun_out
: > $out
run() { echo "=== $*" >> $out; env "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run BOBE_X=1 timeout 600 python tools/factor_ab.py check
run BOBE_LEAF=1 timeout 600 python tools/factor_ab.py check
run BOBE_FACTOR_PW=1 timeout 600 python tools/factor_ab.py check
run BOBE_LOOKAHEAD_MAX=0 BOBE_FACTOR_PW=3 timeout 600 python tools/factor_ab.py check
run BOBE_FACTOR=0 timeout 600 python tools/factor_ab.py time
run BOBE_LEAF=1 timeout 600 python tools/factor_ab.py time
for pw in 1 2 4 8 64; do
  run BOBE_FACTOR_PW=$pw timeout 600 python tools/factor_ab.py time
done
run BOBE_LOOKAHEAD_MAX=0 timeout 600 python tools/factor_ab.py time
run BOBE_LOOKAHEAD_MAX=0 BOBE_FACTOR_PW=64 timeout 600 python tools/factor_ab.py time
run BOBE_LOOKAHEAD_MAX=64 BOBE_MLL_STREAMS=1 timeout 600 python tools/factor_ab.py time
run BOBE_LOOKAHEAD_MAX=64 BOBE_MLL_MIN_PER_STREAM=8 timeout 600 python tools/factor_ab.py time
