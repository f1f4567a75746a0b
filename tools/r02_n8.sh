#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value', j['value'], 'e2e', j['e2e']['value'], 'strong', j['strong']['value'], j['strong']['ms_per_step'], 'secondary', j['secondary']['value'], j['secondary']['ms_per_round'], j['secondary']['frac_of_fp64_peak'], 'wipv', j['wipv']['ms_per_call'], j['wipv']['frac_of_fp64_peak'])
PY
tail -2 gpurun_out/r02_bench_n$N.err
