# L2 persistence experiments: set-aside (NAFB_L2_PERSIST_MB) x policy bits
set -x
mkdir -p gpurun_out
run() {
  tag=$1; bits=$2; mb=$3
  NAFB_L2_PERSIST_MB=$mb NAFB_DEBUG_SKIP=$bits timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 200 --warmup 20 > gpurun_out/r3b_bench_$tag.json 2> gpurun_out/r3b_bench_$tag.err; echo "bench $tag rc $?"
  grep "nafb. L2" gpurun_out/r3b_bench_$tag.err | head -1
  python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('CFG', sys.argv[2], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['final_loss'], {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r3b_bench_$tag.json $tag
  NAFB_L2_PERSIST_MB=$mb NAFB_DEBUG_SKIP=$bits ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_density|k_adam" -c 60 --csv --log-file gpurun_out/r3b_live_$tag.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile-steps 3 > gpurun_out/ncu_live_$tag.log 2>&1
  python - $tag <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f"gpurun_out/r3b_live_{sys.argv[1]}.csv")) if len(r) > 10]
hdr = rows[0]; ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
by = {}
for r in rows[1:]:
    by.setdefault((int(r[iid]), r[ik].split("(")[0].split("::")[-1][:18]), {})[r[im]] = float(r[iv])
last = sorted(by.items())[-3:]
tot = 0
for (i, k), m in last:
    print("   ", k, "read MB", round(m['dram__bytes_read.sum'] / 1e6, 1), "write MB", round(m['dram__bytes_write.sum'] / 1e6, 1), "us", round(m['gpu__time_duration.sum'] / 1e3, 1))
    tot += m['dram__bytes_read.sum'] + m['dram__bytes_write.sum']
print("    step total MB", round(tot / 1e6, 1))
PY
}
run p0 0 0
run pmax_p 0 1000
run pmax_pg $(( (1<<22) + (1<<24) )) 1000
run pmax_all $(( 64 + 128 + (1<<22) + (1<<23) + (1<<24) )) 1000
run p64_pg $(( (1<<22) + (1<<24) )) 64
run p32_pg $(( (1<<22) + (1<<24) )) 32
