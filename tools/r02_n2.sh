#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -s > gpurun_out/r02_dist_check_n2.log 2>&1; echo "dist rc=$?"; tail -5 gpurun_out/r02_dist_check_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value', j['value'], 'e2e', j['e2e']['value'], 'strong', j['strong']['value'], 'secondary', j['secondary']['value'], j['secondary']['ms_per_round'], 'wipv', j['wipv']['ms_per_call'])
PY
tail -3 gpurun_out/r02_bench_n2.err
