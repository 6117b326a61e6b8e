# native op with pair merging + warp aggregation: parity tests, timing against the reference's extension
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2f_parity.log 2>&1; echo "parity rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2f_parity.log | tail -6
timeout 300 python scripts/native_op_time.py 2>&1 | grep -v "Warning\|custom_" | tail -4 | tee gpurun_out/r2f_native_op_time.log
