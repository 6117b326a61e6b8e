set -x
mkdir -p gpurun_out
NAFB_BWD_SW=8 NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2h_stamps_sw8.log 2>&1
grep -v "occupancy" gpurun_out/r2h_stamps_sw8.log | head -14
