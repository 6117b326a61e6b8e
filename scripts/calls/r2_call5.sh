# round 2, call 5: why one CTA per SM? + fixed tests
set -x
mkdir -p gpurun_out
NAFB_BWD_SW=4 NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2e_stamps_sw4.log 2>&1
grep "nafb" gpurun_out/r2e_stamps_sw4.log | head -50
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2e_tests.log | tail -30
