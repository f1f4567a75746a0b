#!/bin/bash
for sv in 1 0 1 0; do BOBE_MLL_SIDE_VECTORS=$sv BOBE_MLL_GRAPH=0 python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n' | grep "R=8\|R=16"; done
for sv in 1 0; do BOBE_MLL_SIDE_VECTORS=$sv python tools/shard_time.py 2>&1 | tail -1; done
