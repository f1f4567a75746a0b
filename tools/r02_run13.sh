#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_run13.log; : > $out
run() { env "$@" >> $out 2>&1; }
run BOBE_X=1 python tools/r64_time.py
run BOBE_FACTOR=0 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=0 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=0 BOBE_MLL_OWN_STREAM=0 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=0 BOBE_MLL_OWN_STREAM=0 BOBE_TINY_MAX_TILES=0 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 BOBE_MLL_OWN_STREAM=0 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 BOBE_FACTOR_PW=8 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 BOBE_FACTOR_PW=2 python tools/r64_time.py
run BOBE_LOOKAHEAD_MAX=8 BOBE_MLL_STREAMS=8 python tools/r64_time.py
cat $out
