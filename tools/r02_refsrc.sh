#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "reference_source" > gpurun_out/r02_refsrc_tests.log 2>&1; echo "pytest rc=$?"; grep "reference source\|reference fit\|passed\|failed\|Error" gpurun_out/r02_refsrc_tests.log | head -20
