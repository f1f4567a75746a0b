#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "shared_panel or config_d or kernel_variants or small_training" > gpurun_out/r02_share_tests2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_share_tests2.log
