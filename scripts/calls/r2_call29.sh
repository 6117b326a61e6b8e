# L2 residency experiments (NAFB_DEBUG_SKIP bits: 64 discard stash after load, 128 stash evict-first, 1<<22 reds evict-last,
# 1<<23 gathers evict-last, 1<<24 Adam keeps the gradient evict-last): bench ms/step + live DRAM traffic per kernel
set -x
mkdir -p gpurun_out
run() {
  tag=$1; bits=$2
  NAFB_DEBUG_SKIP=$bits timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 200 --warmup 20 > gpurun_out/r3a_bench_$tag.json 2> gpurun_out/r3a_bench_$tag.err; echo "bench $tag rc $?"
  python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('CFG', sys.argv[2], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['final_loss'], {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r3a_bench_$tag.json $tag
  NAFB_DEBUG_SKIP=$bits ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_density|k_adam" -c 60 --csv --log-file gpurun_out/r3a_live_$tag.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile-steps 3 > gpurun_out/ncu_live_$tag.log 2>&1
  python - $tag <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f"gpurun_out/r3a_live_{sys.argv[1]}.csv")) if len(r) > 10]
hdr = rows[0]; ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
by = {}
for r in rows[1:]:
    by.setdefault((int(r[iid]), r[ik].split("(")[0].split("::")[-1][:18]), {})[r[im]] = float(r[iv])
last = sorted(by.items())[-6:]
tot_r = tot_w = 0
for (i, k), m in last[-3:]:
    print("   ", k, "read MB", round(m['dram__bytes_read.sum'] / 1e6, 1), "write MB", round(m['dram__bytes_write.sum'] / 1e6, 1), "us", round(m['gpu__time_duration.sum'] / 1e3, 1))
    tot_r += m['dram__bytes_read.sum']; tot_w += m['dram__bytes_write.sum']
print("    step total MB", round((tot_r + tot_w) / 1e6, 1))
PY
}
run base 0
run adamg $((1<<24))
run stash $((128 + (1<<24)))
run discard $((64 + 128 + (1<<24)))
run reds $((64 + 128 + (1<<22) + (1<<24)))
run all $((64 + 128 + (1<<22) + (1<<23) + (1<<24)))
