# schedules of the predict step: time (plain run) and DRAM bytes per kernel (ncu, metrics only)
for cfg in "1 3" "2 1" "4 1" "4 2" "8 1"; do
  set -- $cfg
  export BOBE_TRMM_SPLIT=$1 BOBE_KCHUNKS=$2
  python tools/split_probe.py 5 2>&1 | tail -1
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"trmm_sumsq|kmat_kernel|trmm_finish" -s 40 -c 40 --csv --log-file gpurun_out/split_$1_$2.csv python tools/split_probe.py 1 > /dev/null 2>&1
done
