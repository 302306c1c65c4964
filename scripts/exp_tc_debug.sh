for d in 0 1 2 4 6 7; do echo "== debug $d"; EFGB_TC_DEBUG=$d python scripts/bench_conv.py fp32x3 2>&1 | grep "subm fwd"; done
