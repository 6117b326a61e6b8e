# ncu capture of the voxel-query kernel (128^3 launch of scripts/voxel_time.py); TAG names the round / state of the code
set -x
TAG=${TAG:-r2}
timeout 200 python scripts/voxel_time.py > gpurun_out/voxel_plain.log 2>&1 || exit 1
cat gpurun_out/voxel_plain.log
ncu --set full --clock-control none --import-source on -k regex:k_density_fwd_tc -s 3 -c 1 -f -o gpurun_out/${TAG}_voxel_query_128 python scripts/voxel_time.py > gpurun_out/ncu_voxel.log 2>&1
echo "voxel exit $?"
