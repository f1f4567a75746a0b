# ncu capture of the DMMA GEMM launches of one 64-restart log-ML+grad round (single chain), sections only
set -x
export BOBE_MLL_STREAMS=1
python tools/prof_run.py mll 64 > gpurun_out/plain_mll.log 2>&1 &&
ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section WarpStateStats \
    --section LaunchStats --section Occupancy --section SchedulerStats --clock-control none \
    -k regex:gemm_nt_kernel -s 181 -c 181 -o gpurun_out/prof_gemm_mll -f python tools/prof_run.py mll 64 > gpurun_out/ncu_gemm_mll.log 2>&1
tail -n 2 gpurun_out/ncu_gemm_mll.log
