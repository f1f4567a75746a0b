#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run18.log; : > $out
BOBE_X=1 python tools/shard_time.py >> $out 2>&1
BOBE_MLL_GRAPH=0 python tools/shard_time.py >> $out 2>&1
BOBE_MLL_GRAPH=0 BOBE_FACTOR=0 python tools/shard_time.py >> $out 2>&1
BOBE_MLL_GRAPH=0 BOBE_MLL_STREAMS=1 python tools/shard_time.py >> $out 2>&1
cat $out
