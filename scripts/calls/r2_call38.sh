# voxel tile shape / lane order experiment, part 2 (NAFB_DEBUG_SKIP bits 26-28): 0 = 4x4x8 k-fastest; i-fastest: 1 8x4x4, 2 8x2x8,
# 3 16x4x2, 4 16x2x4, 5 32x2x2, 6 32x4x1, 7 128x1x1
set -x
for m in 0 1 2 3 4 5 6 7; do
  echo "=== vox_mode $m"
  NAFB_DEBUG_SKIP=$((m << 26)) timeout 300 python scripts/voxel_time.py 2>&1 | grep "voxel query"
done
