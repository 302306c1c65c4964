for d in ${DBG:-0 7 15 23 31 63}; do echo "== debug $d"; EFGB_TC_DEBUG=$d python scripts/bench_conv.py fp32x3 2>&1 | grep "subm fwd"; done
