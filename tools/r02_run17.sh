#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run17.log; : > $out
for la in 8 16; do for s in 1 2 3 4; do BOBE_MLL_GRAPH=0 BOBE_MLL_MIN_PER_STREAM=1 BOBE_MLL_STREAMS=$s BOBE_LOOKAHEAD_MAX=$la python tools/sub_sweep.py >> $out 2>&1; done; done
cat $out
