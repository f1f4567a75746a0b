#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "reference_source or reference_autodiff" > gpurun_out/r02_refsrc_tests.log 2>&1; echo "pytest rc=$?"; grep "reference source\|reference fit\|reference flows\|reference autodiff\|passed\|failed\|Error" gpurun_out/r02_refsrc_tests.log | head -20
