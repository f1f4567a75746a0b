#!/bin/bash
# round-2 profile captures (each ncu pass only after the same command exited 0 without ncu)
mkdir -p gpurun_out
export BOBE_MLL_GRAPH=0
CMD="python bench.py --steps 1 --warmup 3 --m-per-gpu 151552 --skip-cpu-baseline --skip-extras"
$CMD > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_bench.log 2>&1
tail -n 1 gpurun_out/r02_ncu_bench.log
python tools/prof_run.py mll 8 > gpurun_out/r02_plain_mll8.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_mll_r8.csv python tools/prof_run.py mll 8 > gpurun_out/r02_ncu_mll8.log 2>&1
python tools/prof_run.py mll 64 > gpurun_out/r02_plain_mll64.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_mll_r64.csv python tools/prof_run.py mll 64 > gpurun_out/r02_ncu_mll64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_leaf128 -s 20 -c 1 -o gpurun_out/r02_prof_leaf -f python tools/prof_run.py mll 8 > gpurun_out/r02_ncu_leaf.log 2>&1
tail -n 2 gpurun_out/r02_ncu_leaf.log
ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_kernel.*32, *64" -s 40 -c 1 -o gpurun_out/r02_prof_gemm_tiny -f python tools/prof_run.py mll 8 > gpurun_out/r02_ncu_gemm_tiny.log 2>&1
tail -n 2 gpurun_out/r02_ncu_gemm_tiny.log
ls -la gpurun_out/r02_prof_*.ncu-rep
