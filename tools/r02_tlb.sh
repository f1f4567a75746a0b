#!/bin/bash
mkdir -p gpurun_out
python tools/timeline.py predict 500 4 100000 rbf > gpurun_out/r02_tl_predict_b.txt 2>&1
python tools/timeline.py predict 1500 27 151552 rbf > gpurun_out/r02_tl_predict_d.txt 2>&1
grep -v Warn gpurun_out/r02_tl_predict_b.txt | head -40
