#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02r.csv python bench.py --steps 1 --warmup 3 --quick --no-overlap > gpurun_out/bench_under_ncu_r02r.log 2>&1
python - <<'PY'
import csv,re
lines=[l for l in open("gpurun_out/launches_r02r.csv") if not l.startswith("==")]
rows=[]
for r in csv.DictReader(lines):
    if r.get("Metric Name")!="gpu__time_duration.sum": continue
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    if u=="ns": v/=1e3
    elif u=="ms": v*=1e3
    rows.append((r["Kernel Name"], v))
last=max(i for i,(n,_) in enumerate(rows) if "fps_regs_kernel<512, 24>" in n)
out=[]
for n,v in rows[last:last+110]:
    n=re.sub(r"pvn3d::<unnamed>::","",n); n=re.sub(r"\(.*","",n)
    out.append(f"{v:8.1f}  {n[:50]}")
    if "transpose" in n: break
print("\n".join(out[-28:]))
PY
