set -x
TAG=${TAG:-r1c}
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-steps 3"
timeout 300 $CMD > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_density_bwd_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_density_bwd_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_density_fwd_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_density_fwd_tc $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full2 exit $?"
