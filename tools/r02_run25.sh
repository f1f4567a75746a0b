#!/bin/bash
python tools/factor_ab.py check 2>&1 | tail -2
BOBE_LOOKAHEAD_MAX=0 python tools/factor_ab.py check 2>&1 | tail -1
python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n'
BOBE_FACTOR_LIVE=0 python tools/r64_time.py 2>&1 | tail -1 | tr '|' '\n'
python tools/factor_ab.py time 2>&1 | grep "factorize"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
