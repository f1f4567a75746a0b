set -x
python tools/prof_run.py predict > gpurun_out/plain_predict.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:kmat_kernel -s 2 -c 1 -o gpurun_out/prof_kmat_v0 -f python tools/prof_run.py predict > gpurun_out/ncu_kmat.log 2>&1
python tools/prof_run.py mll 64 > gpurun_out/plain_mll.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:leaf128 -s 20 -c 1 -o gpurun_out/prof_leaf128_v0 -f python tools/prof_run.py mll 64 > gpurun_out/ncu_leaf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mll_grad_tile -s 1 -c 1 -o gpurun_out/prof_gradtile_v0 -f python tools/prof_run.py mll 64 > gpurun_out/ncu_gradtile.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_kernel.*128, 128" -s 6 -c 3 -o gpurun_out/prof_gemmbig_v0 -f python tools/prof_run.py mll 64 > gpurun_out/ncu_gemmbig.log 2>&1
tail -2 gpurun_out/ncu_*.log
