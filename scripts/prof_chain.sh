set -x
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-steps 3"
NAFB_DEBUG_SKIP=1 ncu --set full --clock-control none --import-source on -k regex:k_density_bwd_tc -s 2 -c 1 -f -o gpurun_out/r1d_bwd_noscatter $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
