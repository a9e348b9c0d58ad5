#!/bin/bash
# tools/ncu_launches_ns.sh -- launch list (gpu__time_duration.sum, no clock control) of the benchmark command itself
# (`python bench.py`, north-star batch) for profiles/: the kernel's share of the step must agree with bench.py's own timing.
O=gpurun_out
mkdir -p $O
ncu --target-processes all python -c "import os; print(sorted(k for k in os.environ if 'INJECT' in k or 'NSIGHT' in k or k.startswith('NV_')))" > $O/ncu_env.txt 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_ns.csv \
  python bench.py --steps 2 --warmup 3 --no-verify --cpu-sample 64 > $O/ncu_launches_ns_bench.log 2>&1
tail -n 3 $O/ncu_env.txt; grep -c lzb_ $O/ncu_launches_ns.csv; python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_launches_ns.csv")) if len(r)>14 and r[0].isdigit()]
for r in rows: print(r[4], r[8], r[7], "%.3f ms"%(float(r[-1])/1e6))
PY
