python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/quick_bench.py 2>&1 | head -12
for st in 2 4 8; do echo "== streams $st"; BOBE_MLL_STREAMS=$st python tools/quick_bench.py 2>&1 | grep -E "^mll\+"; done
