set -x
mkdir -p gpurun_out
NAFB_FWD_STAMPS=1 NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2l_fwd_stamps.log 2>&1
grep "loop-exit\|histogram\|CUDA events" gpurun_out/r2l_fwd_stamps.log | tail -4
grep -A7 "kernel milestones" gpurun_out/r2l_fwd_stamps.log
