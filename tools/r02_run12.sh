#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run12.log; : > $out
run() { echo "=== $*" >> $out; env "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run BOBE_GEMM_TMA=2 BOBE_GEMM_TMA_MIN_TILES=1 BOBE_LOOKAHEAD_MAX=0 timeout 600 python tools/factor_ab.py check
run BOBE_GEMM_TMA=2 BOBE_GEMM_TMA_MIN_TILES=1 BOBE_LOOKAHEAD_MAX=0 timeout 600 python tools/factor_ab.py time
run BOBE_GEMM_TMA=2 BOBE_GEMM_TMA_MIN_TILES=1 BOBE_LOOKAHEAD_MAX=0 BOBE_MLL_STREAMS=1 timeout 600 python tools/factor_ab.py time
run BOBE_GEMM_TMA=2 BOBE_GEMM_TMA_MIN_TILES=1 BOBE_LOOKAHEAD_MAX=0 BOBE_MLL_STREAMS=2 timeout 600 python tools/factor_ab.py time
run BOBE_GEMM_TMA=1 BOBE_GEMM_TMA_MIN_TILES=1 BOBE_LOOKAHEAD_MAX=0 BOBE_MLL_STREAMS=2 timeout 600 python tools/factor_ab.py time
run BOBE_LOOKAHEAD_MAX=0 BOBE_MLL_STREAMS=2 timeout 600 python tools/factor_ab.py time
grep -v "^n=" $out | grep "===\|R=16\|R=64\|check ok\|FAILED\|rror"
