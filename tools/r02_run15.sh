#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run15.log; : > $out
run() { env "$@" >> $out 2>&1; }
run BOBE_X=1 python tools/r64_time.py
run BOBE_MLL_GRAPH=0 python tools/r64_time.py
cat $out | tr '|' '\n'
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests2.log 2>&1; tail -4 gpurun_out/r02_gpu_tests2.log
