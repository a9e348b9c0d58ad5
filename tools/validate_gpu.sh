#!/bin/bash
# tools/validate_gpu.sh -- one 1-GPU box call: GPU tests, smoke(), differential fuzz soak of both entry points, compute-sanitizer,
# the benchmark line (3 steps).  Output under gpurun_out/lat3_*.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/lat3_pytest.txt
timeout 300 python __graft_entry__.py smoke > $O/lat3_smoke.txt 2>&1
for seed in 221 222; do timeout 300 python tools/fuzz_soak.py --backend gpu --rounds 30 --seed $seed; done > $O/lat3_fuzz.txt 2>&1
timeout 300 python tools/fuzz_soak.py --backend gpu-device --rounds 30 --seed 223 >> $O/lat3_fuzz.txt 2>&1
timeout 300 python tools/fuzz_soak.py --backend gpu --rounds 20 --seed 224 --cases >> $O/lat3_fuzz.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/lat3_memcheck.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/lat3_racecheck.txt 2>&1
( LZB_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 3 ) > $O/lat3_bench_ns.json 2> $O/lat3_bench_ns.err
cat $O/lat3_pytest.txt $O/lat3_smoke.txt $O/lat3_fuzz.txt; tail -n 4 $O/lat3_memcheck.txt; tail -n 4 $O/lat3_racecheck.txt; cut -c1-200 $O/lat3_bench_ns.json
