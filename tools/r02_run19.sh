#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run19.log; : > $out
for s in 1 2 3 4; do BOBE_MLL_GRAPH=0 BOBE_MLL_MIN_PER_STREAM=1 BOBE_MLL_STREAMS=$s python tools/sub_sweep.py >> $out 2>&1; done
BOBE_MLL_GRAPH=0 BOBE_MLL_MIN_PER_STREAM=1 BOBE_MLL_STREAMS=1 BOBE_LOOKAHEAD_MAX=16 python tools/sub_sweep.py >> $out 2>&1
BOBE_MLL_GRAPH=0 BOBE_MLL_MIN_PER_STREAM=1 BOBE_MLL_STREAMS=2 BOBE_LOOKAHEAD_MAX=16 python tools/sub_sweep.py >> $out 2>&1
BOBE_MLL_GRAPH=0 BOBE_MLL_MIN_PER_STREAM=1 BOBE_MLL_STREAMS=1 BOBE_FACTOR=0 python tools/sub_sweep.py >> $out 2>&1
BOBE_MLL_GRAPH=0 BOBE_MLL_MIN_PER_STREAM=4 BOBE_MLL_STREAMS=4 BOBE_FACTOR=0 python tools/sub_sweep.py >> $out 2>&1
cat $out
