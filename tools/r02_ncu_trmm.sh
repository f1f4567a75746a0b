#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --skip-cpu-baseline --skip-extras > gpurun_out/r02_ncu_trmm_plain.json 2>/dev/null; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trmm_sumsq_tma --launch-skip 10 --launch-count 1 -o gpurun_out/r02_trmm_full -f python bench.py --steps 1 --warmup 1 --m-per-gpu 151552 --skip-cpu-baseline --skip-extras > gpurun_out/r02_ncu_trmm.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02_trmm_full.ncu-rep
