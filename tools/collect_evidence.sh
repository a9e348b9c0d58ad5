#!/bin/bash
# tools/collect_evidence.sh TAG -- one 1-GPU gpurun call's worth of round evidence into gpurun_out/ (copied to profiles/ afterwards):
# GPU tests, the benchmark line as the driver runs it, the reference arm, the informational configs, the host-API timeline,
# the ncu launch list and one --set full capture of K1, a differential fuzz soak of both entry points.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
( LZB_TRACE=1 timeout 1500 python bench.py --steps 20 --warmup 5 ) > $O/${TAG}_bench_ns_1gpu.json 2> $O/${TAG}_bench_ns_1gpu.err
for c in c2 c3 c5 c6 c4; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --cpu-sample 1024 > $O/${TAG}_bench_$c.json 2> $O/${TAG}_bench_$c.err
done
timeout 600 python tools/kbench.py --config c2 --streams 1024 --steps 7 lzma_rs_b200/liblzma_b200.so > $O/${TAG}_kbench_c2_1024.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --config c2 --steps 2 --warmup 3 --no-verify --cpu-sample 64 > $O/${TAG}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lzb_decode_kernel -s 3 -c 1 -f -o $O/${TAG}_k1_full \
  python bench.py --config c2 --steps 2 --warmup 3 --no-verify --cpu-sample 64 > $O/${TAG}_k1_full.log 2>&1
for seed in 211 212; do timeout 300 python tools/fuzz_soak.py --backend gpu --rounds 30 --seed $seed; done > $O/${TAG}_fuzz_gpu.txt 2>&1
timeout 300 python tools/fuzz_soak.py --backend gpu-device --rounds 30 --seed 213 >> $O/${TAG}_fuzz_gpu.txt 2>&1
timeout 300 python tools/fuzz_soak.py --backend gpu --rounds 20 --seed 214 --cases >> $O/${TAG}_fuzz_gpu.txt 2>&1
ls -la $O | tail -30
