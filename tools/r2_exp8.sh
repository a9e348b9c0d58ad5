#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > $O/r2_exp8_pytest.txt
( LZB_TRACE=1 timeout 1500 python bench.py --steps 3 --warmup 3 ) > $O/r2_exp8_ns_full.json 2> $O/r2_exp8_ns_full.err
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:lzb_decode -s 3 -c 1 --csv --log-file $O/r2_exp8_ns_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no-verify --cpu-sample 8 > $O/r2_exp8_ns_traffic.log 2>&1
cat $O/r2_exp8_pytest.txt; grep lzb_trace $O/r2_exp8_ns_full.err | tail -2; cat $O/r2_exp8_ns_full.json | cut -c1-400; tail -5 $O/r2_exp8_ns_traffic.csv
