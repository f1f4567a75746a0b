#!/bin/bash
mkdir -p gpurun_out
BOBE_PDL=0 python tools/timeline.py factor 2000 > gpurun_out/r02_tl3_factor_nopdl.txt 2>&1
out=gpurun_out/r02_run3.log; : > $out
run() { echo "=== $*" >> $out; env "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run BOBE_TINY_MAX_TILES=296 timeout 600 python tools/factor_ab.py time
run BOBE_TINY_MAX_TILES=592 timeout 600 python tools/factor_ab.py time
run BOBE_PDL=0 timeout 600 python tools/factor_ab.py time
grep -v "^n=" $out | grep "===\|factorize n=2000\|R=8\|R=1:"
