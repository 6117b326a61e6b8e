# high-priority draw stream: sampled leg; live (warm-cache) DRAM traffic of the step's kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_select.py -m gpu -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2z_tests.log | tail -8
timeout 900 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "bench rc $?"
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']/1e6,1), 'sampled', round(d['sampled']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r2z_bench_n1.json
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile-steps 3"
ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_density|k_adam" -c 60 --csv --log-file gpurun_out/r2z_live_traffic.csv $CMD > gpurun_out/ncu_live.log 2>&1
echo "live traffic exit $?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2z_live_traffic.csv")) if len(r) > 10]
hdr = rows[0]; ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iid = hdr.index("ID")
by = {}
for r in rows[1:]:
    by.setdefault((int(r[iid]), r[ik].split("(")[0][:40]), {})[r[im]] = r[iv]
for (i, k), m in sorted(by.items())[-12:]:
    print(i, k, m)
PY
