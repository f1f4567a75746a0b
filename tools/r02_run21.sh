#!/bin/bash
BOBE_MLL_GRAPH=0 python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n' | grep "R=8"
BOBE_MLL_GRAPH=0 BOBE_MLL_STREAMS=2 BOBE_MLL_MIN_PER_STREAM=1 python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n' | grep "R=8"
BOBE_MLL_GRAPH=0 BOBE_MLL_STREAMS=4 BOBE_MLL_MIN_PER_STREAM=4 python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n' | grep "R=8"
BOBE_MLL_GRAPH=0 BOBE_MLL_STREAMS=1 python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n' | grep "R=8"
BOBE_MLL_GRAPH=0 python tools/factor_ab.py time 2>&1 | grep "R=8\|factorize n=2000"
