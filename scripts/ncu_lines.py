#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).

  python scripts/ncu_lines.py gpurun_out/X.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

src = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.check_output(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "cuda,sass"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = "?", None
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 10 and r[0].strip().isdigit():
        d = dict(zip(hdr, r))
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        num = lambda v: int(v) if v.strip().isdigit() else 0
        lines.append((num(d["# Samples"]), cur_file, int(r[0]), r[1].strip()[:90], num(d["Instructions Executed"]), stalls))
total = sum(l[0] for l in lines)
print(f"# {src}: {total} samples")
tot_st = {}
for l in lines:
    for k, v in l[5].items():
        tot_st[k] = tot_st.get(k, 0) + v
print("# stall totals:", ", ".join(f"{k}={100*v/total:.1f}%" for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:10]))
for s, f, ln, text, inst, st in sorted(lines, key=lambda x: -x[0])[:top]:
    main = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/total:5.1f}%  {f}:{ln:<4d} inst={inst:<9d} {text}   [{main}]")
