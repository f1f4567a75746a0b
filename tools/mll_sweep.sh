#!/bin/bash
# sweep of the two scheduling knobs of the batched log-ML+grad call
for st in 4 8; do for th in 74 148 296; do
  echo "== streams $st small_below $th"
  BOBE_MLL_STREAMS=$st BOBE_SMALL_TILE_CTAS=$th python tools/quick_bench.py 2>&1 | grep -E "mll\+"
done; done
