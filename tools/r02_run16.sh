#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run16.log; : > $out
run() { env "$@" >> $out 2>&1; }
run BOBE_X=1 python tools/bench_mll_time.py
run BOBE_FACTOR=0 python tools/bench_mll_time.py
run BOBE_MLL_GRAPH=0 python tools/bench_mll_time.py
python tools/fit_bench.py 2000 64 >> $out 2>&1
cat $out | grep -v Warning
