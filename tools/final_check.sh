#!/bin/bash
# tools/final_check.sh -- last 1-GPU call of the round: GPU tests and smoke() on the final build, the benchmark line as the
# driver runs it, one --set full capture of K1 on a north-star shard (8 192 mixed-size streams).
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 4 > $O/final_pytest.txt
timeout 300 python __graft_entry__.py smoke > $O/final_smoke.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lzb_decode_sched_kernel -s 3 -c 1 -f -o $O/final_k1_ns8192 \
  python bench.py --streams 8192 --distinct 1024 --steps 2 --warmup 3 --no-verify --cpu-sample 64 > $O/final_k1_ns8192.log 2>&1
( timeout 1200 python bench.py --steps 20 --warmup 5 ) > $O/final_bench_ns_1gpu.json 2> $O/final_bench_ns_1gpu.err
cat $O/final_pytest.txt $O/final_smoke.txt; tail -n 3 $O/final_k1_ns8192.log; cut -c1-200 $O/final_bench_ns_1gpu.json
