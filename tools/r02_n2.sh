#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -s > gpurun_out/r02_dist_check_n2.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r02_dist_check_n2.log
bash tools/r02_n8.sh 2
