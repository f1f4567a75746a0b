#!/bin/bash
mkdir -p gpurun_out
BOBE_MLL_MIN_PER_STREAM=8 python tools/timeline.py mll 8 > gpurun_out/r02_tl5_mll8_s1.txt 2>&1
python tools/timeline.py mll 8 > gpurun_out/r02_tl5_mll8_s2.txt 2>&1
head -5 gpurun_out/r02_tl5_mll8_s1.txt
