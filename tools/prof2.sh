set -x
CMD="python bench.py --steps 1 --warmup 3 --m-per-gpu 151552 --skip-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
python tools/prof_run.py predict > gpurun_out/plain_predict.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"trmm_sumsq|kmat_kernel" -s 4 -c 2 -o gpurun_out/prof_predict_final -f python tools/prof_run.py predict > gpurun_out/ncu_predict.log 2>&1
tail -n 2 gpurun_out/ncu_bench.log; tail -n 2 gpurun_out/ncu_predict.log
