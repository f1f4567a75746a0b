#!/bin/bash
mkdir -p gpurun_out
python tools/timeline.py factor 2000 > gpurun_out/r02_tl_factor.txt 2>&1
python tools/timeline.py mll 8 > gpurun_out/r02_tl_mll8.txt 2>&1
python tools/timeline.py mll 64 > gpurun_out/r02_tl_mll64.txt 2>&1
head -3 gpurun_out/r02_tl_factor.txt
