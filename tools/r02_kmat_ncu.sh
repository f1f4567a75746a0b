#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmat_kernel --launch-skip 0 --launch-count 1 -o gpurun_out/r02_kmat_mll -f python tools/kmat_probe.py mll > gpurun_out/r02_kmat_mll.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mll_grad_tile --launch-skip 3 --launch-count 1 -o gpurun_out/r02_gradtile -f python tools/kmat_probe.py mll > gpurun_out/r02_gradtile.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
