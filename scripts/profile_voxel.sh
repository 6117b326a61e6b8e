set -x
timeout 200 python scripts/voxel_time.py > gpurun_out/voxel_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_density_fwd_tc -s 1 -c 1 -f -o gpurun_out/r1f_voxel_fwd python scripts/voxel_time.py > gpurun_out/ncu_voxel.log 2>&1
echo "voxel exit $?"
