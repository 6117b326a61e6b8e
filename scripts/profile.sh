# ncu captures of the training step's kernels (one GPU; run under gpurun).  TAG names the round / state of the code.
set -x
TAG=${TAG:-r2}
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile-steps 3"
timeout 300 $CMD > gpurun_out/prof_plain.log 2>&1 || exit 1
for k in k_density_bwd_ws k_density_fwd_tc k_adam_dev; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/${TAG}_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "$k exit $?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_density|k_adam|k_mse|k_reduce|k_draw" -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch-list exit $?"
timeout 200 python scripts/voxel_time.py > gpurun_out/voxel_plain.log 2>&1 || exit 1
cat gpurun_out/voxel_plain.log
ncu --set full --clock-control none --import-source on -k regex:k_density_fwd_tc -s 3 -c 1 -f -o gpurun_out/${TAG}_voxel_query_128 python scripts/voxel_time.py > gpurun_out/ncu_voxel.log 2>&1
echo "voxel exit $?"
