set -x
mkdir -p gpurun_out
for v in "NAFB_CARVEOUT=-1" "NAFB_CARVEOUT=100" "NAFB_CARVEOUT=75"; do
  echo "== variant [$v]"
  env $v timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra --profile-steps 10 2>gpurun_out/r2p_bench_err.log > gpurun_out/r2p_bench.json
  python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r2p_bench.json
done
