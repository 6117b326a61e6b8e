# voxel tile shape / lane order experiment (NAFB_DEBUG_SKIP bits 26-27): 0 = 4x4x8 k-fastest, 1 = 8x4x4 i-fastest, 2 = 8x2x8 i-fastest
set -x
mkdir -p gpurun_out
for m in 0 1 2; do
  echo "=== vox_mode $m"
  NAFB_DEBUG_SKIP=$((m << 26)) timeout 300 python scripts/voxel_time.py 2>&1 | grep "voxel query"
done
for m in 1 2; do
  NAFB_DEBUG_SKIP=$((m << 26)) timeout 600 python -m pytest tests -m gpu -q -k "voxel" 2>&1 | tail -2
done
