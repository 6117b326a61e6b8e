set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_exchange.py -q -x -k "busy" > gpurun_out/r2s_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2s_tests.log | tail -8
