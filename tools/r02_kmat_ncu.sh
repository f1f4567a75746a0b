#!/bin/bash
mkdir -p gpurun_out
python tools/kmat_probe.py mll && python tools/kmat_probe.py b || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmat_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r02_kmat_mll -f python tools/kmat_probe.py mll > gpurun_out/r02_kmat_mll.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmat_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_kmat_b -f python tools/kmat_probe.py b > gpurun_out/r02_kmat_b.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
