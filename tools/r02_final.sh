#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gpu_tests_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
python tools/flow_bench.py > gpurun_out/r02_flow_final.txt 2>&1
python tools/factor_ab.py time > gpurun_out/r02_factor_time_final.txt 2>&1
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'secondary', j['secondary']['value'], j['secondary']['ms_per_round'], j['secondary']['frac_of_fp64_peak'], 'wipv', j['wipv']['ms_per_call'], 'cpu', j['cpu_baseline']['value'])
r=json.loads(open('gpurun_out/r02_bench_reference.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['secondary']['value'], r['cpu_baseline']['sample'][:200])
PY
grep "get_next_point\|->" gpurun_out/r02_flow_final.txt | head -20
grep "factorize\|mll" gpurun_out/r02_factor_time_final.txt
