#!/usr/bin/env bash
# per-kernel durations of the four fused ball-query+group launches (GPU box)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/qg_launches.csv \
    python tools/qg_roofline.py > gpurun_out/qg_roofline_under_ncu.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/qg_launches.csv') if not l.startswith('==')]
rows=[(r['Kernel Name'][:60], r['Grid Size'], float(r['Metric Value'].replace(',','')), r['Metric Unit']) for r in csv.DictReader(lines)]
for r in rows[-40:]: print(r)
PY
