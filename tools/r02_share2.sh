#!/bin/bash
mkdir -p gpurun_out
run() { timeout 600 python bench.py --skip-extras --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'ms_per_launch', j['roofline']['ms_per_launch'], 'frac', j['roofline']['frac'], 'W', j['clocks']['power_w_max'])
"; }
for rep in 1 2; do
  for lib in old new; do cp tools/_ab/lib$lib.so bobe_b200/lib/libbobe_b200.so; echo "== lib $lib (schedule off)"; run; done
done
cp tools/_ab/libnew.so bobe_b200/lib/libbobe_b200.so
for s in 4 2; do echo "== new, BOBE_TRMM_SHARE=$s"; BOBE_TRMM_SHARE=$s run; done
for s in 1 4; do
  BOBE_TRMM_SHARE=$s timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --cache-control none -k regex:trmm_sumsq_tma --launch-skip 12 --launch-count 6 --csv --log-file gpurun_out/r02_share_ncu_s$s.csv python bench.py --skip-extras --steps 1 --warmup 1 > gpurun_out/r02_share_ncu_s$s.log 2>&1; echo "ncu S=$s rc=$?"
done
python - <<'PY'
import csv
for s in (1, 4):
    rows = [r for r in csv.reader(open(f'gpurun_out/r02_share_ncu_s{s}.csv')) if len(r) > 10]
    hdr = rows[0]; mi, vi, ki, ii = hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Kernel Name'), hdr.index('ID')
    agg = {}
    for r in rows[1:]:
        agg.setdefault(r[ii], {})[r[mi]] = float(r[vi].replace(',', ''))
    for i, m in agg.items():
        print('S', s, 'launch', i, m)
PY
