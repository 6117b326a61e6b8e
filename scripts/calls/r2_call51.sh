# native op: staged-output forward, point-looped backward: parity tests + timing against the reference's extension
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainer.py -m gpu -q -x > gpurun_out/r2g_parity.log 2>&1; echo "parity rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2g_parity.log | tail -6
timeout 300 python scripts/native_op_time.py 2>&1 | grep -v "Warning\|custom_\|_warn_once" | tail -5 | tee gpurun_out/r2g_native_op_time.log
