# how many levels of gather loads in flight per thread pay? (native forward op, thread = point; depth 0 = shipped kernel with 1 ahead)
for d in 0 1 2 3 4 5; do NAFB_DEBUG_SKIP=$((d << 26)) timeout 120 python scripts/native_fwd_time.py 2>&1 | tail -1; done
