# round 2, call 1: baseline sanity + phase stamps of the backward kernel (with / without scatter)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc $?"
NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2a_stamps_scatter.log 2>&1
NAFB_DEBUG_SKIP=33 timeout 120 python scripts/stamps.py > gpurun_out/r2a_stamps_noscatter.log 2>&1
ATTRIB_FLAGS="0 1 2 3" timeout 300 bash scripts/attrib.sh > gpurun_out/r2a_attrib.log 2>&1
tail -3 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_attrib.log
