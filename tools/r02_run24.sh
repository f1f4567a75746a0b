#!/bin/bash
for pw in 2 3 4 6 8 16; do BOBE_FACTOR_PW=$pw BOBE_MLL_GRAPH=0 python tools/r64_time.py 2>&1 | tail -1 | sed 's/; call+fetch[^|]*//g'; done
