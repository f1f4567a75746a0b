#!/bin/bash
BOBE_MLL_GRAPH=0 python tools/factor_ab.py check 2>&1 | tail -1
BOBE_MLL_GRAPH=0 python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n'
BOBE_MLL_GRAPH=0 BOBE_TINY_NARROW=0 python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n'
python tools/shard_time.py 2>&1 | tail -1
python tools/factor_ab.py time 2>&1 | grep "factorize"
