#!/bin/bash
mkdir -p gpurun_out
for g in 1 2 3; do echo "== BOBE_FIT_GROUPS=$g"; BOBE_FIT_GROUPS=$g timeout 600 python tools/fit_bench.py 2000 64 2>&1 | grep -v Warn; done > gpurun_out/r02_fit_groups.txt 2>&1
python tools/flow_bench.py > gpurun_out/r02_flow_pool.txt 2>&1
cat gpurun_out/r02_fit_groups.txt; grep "Surrogate\|fit:" gpurun_out/r02_flow_pool.txt
