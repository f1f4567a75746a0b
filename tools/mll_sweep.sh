#!/bin/bash
for st in 4; do for tile in 1; do
  echo "== streams $st tile $tile"
  BOBE_MLL_STREAMS=$st BOBE_TILE=$tile python tools/quick_bench.py 2>&1 | grep -E "^mll\+"
done; done
