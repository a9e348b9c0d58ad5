#!/bin/bash
# tools/collect_evidence.sh TAG -- one gpurun call's worth of round evidence into gpurun_out/ (copied to profiles/ afterwards):
# GPU tests, the benchmark line, the reference arm, the informational configs, the ncu launch list and one --set full capture.
TAG=${1:-r01}
O=gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/${TAG}_pytest_gpu.log
timeout 400 python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
timeout 400 python bench.py --impl reference > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
for c in c3 c5 c6 c4 ns; do
  timeout 500 python bench.py --config $c --steps 5 --warmup 3 --cpu-sample 1024 > $O/${TAG}_bench_$c.json 2> $O/${TAG}_bench_$c.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-verify --cpu-sample 64 > $O/${TAG}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lzb_decode_kernel -s 3 -c 1 -f -o $O/${TAG}_k1_full \
  python bench.py --steps 2 --warmup 3 --no-verify --cpu-sample 64 > $O/${TAG}_k1_full.log 2>&1
ls -la $O
