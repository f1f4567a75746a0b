#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "predict or golden or var or config_d or config_h or full_size" > gpurun_out/r02_share_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_share_tests.log
for s in 1 0 2 4; do echo "== BOBE_TRMM_SHARE=$s"; BOBE_TRMM_SHARE=$s timeout 600 python bench.py --skip-extras --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'ms_per_launch', j['roofline']['ms_per_launch'], 'frac', j['roofline']['frac'], 'share', j['roofline']['share_of_step'], 'W', j['clocks']['power_w_max'])
"; done
