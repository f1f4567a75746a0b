# the TMA-fed trmm_sumsq kernel: tests, bench, launch list, full capture
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --skip-cpu-baseline > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err
python tools/prof_run.py predict > gpurun_out/plain_predict.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"trmm_sumsq" -s 4 -c 1 -o gpurun_out/prof_trmm_tma -f python tools/prof_run.py predict > gpurun_out/ncu_trmm_tma.log 2>&1
tail -n 2 gpurun_out/ncu_trmm_tma.log
CMD="python bench.py --steps 1 --warmup 3 --m-per-gpu 151552 --skip-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench_tma.csv $CMD > gpurun_out/ncu_bench.log 2>&1
tail -n 1 gpurun_out/ncu_bench.log
