#!/bin/bash
# fixed cost of a launch: ncu durations of tiny launches in the three mappings; ping-pong without the start skew
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_g15
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fa_fwd --csv --log-file gpurun_out/${T}_tiny_launches.csv python tools/tiny_launches.py > gpurun_out/${T}_tiny_launches.log 2>&1; tail -3 gpurun_out/${T}_tiny_launches.log
python - <<'PY'
import csv, re
rows = [r for r in csv.reader(open('gpurun_out/r02_g15_tiny_launches.csv')) if len(r) > 10 and r[0].isdigit()]
for r in rows:
    name = re.sub(r'\(.*', '', r[4]).replace('void fa::', '')
    print(r[0], name[:40], r[7], r[8], r[-2], r[-1])
PY
timeout 300 python tools/sweep_variants.py --timeout 100 --only base,ppnoskew --shapes "16,512,16;16,1024,16;37,128,4" --modes pp --reps 30 --out gpurun_out/${T}_sweep_ppskew.json 2>&1 | tail -8
