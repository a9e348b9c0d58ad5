#!/bin/bash
O=gpurun_out
mkdir -p $O
L=lzma_rs_b200/liblzma_b200.so
for n in 1 148 592 1024 1184; do
  echo "== streams $n" 
  timeout 300 python tools/kbench.py --config c2 --streams $n --steps 7 $L $L@LZB_NO_LAT=1 2>&1 | tail -2
done > $O/lat1_kbench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/lat1_pytest.txt
( LZB_TRACE=1 timeout 600 python bench.py --config c2 --steps 10 --warmup 3 --cpu-sample 256 ) > $O/lat1_bench_c2.json 2> $O/lat1_bench_c2.err
cat $O/lat1_kbench.txt $O/lat1_pytest.txt; tail -2 $O/lat1_bench_c2.err; cut -c1-900 $O/lat1_bench_c2.json
