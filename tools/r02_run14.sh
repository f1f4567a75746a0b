#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run14.log; : > $out
run() { env "$@" >> $out 2>&1; }
run BOBE_X=1 python tools/factor_ab.py check
run BOBE_LOOKAHEAD_MAX=0 python tools/factor_ab.py check
run BOBE_X=1 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 BOBE_FACTOR_PW=8 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 BOBE_FACTOR_PW=2 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 BOBE_MLL_STREAMS=2 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 BOBE_MLL_MIN_PER_STREAM=2 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=4 python tools/r64_time.py
grep "check ok\|FAILED\|R=64" $out
