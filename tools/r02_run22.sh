#!/bin/bash
BOBE_MLL_GRAPH=0 python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n'
python tools/shard_time.py 2>&1 | tail -1
python tools/bench_mll_time.py 2>&1 | tail -1
