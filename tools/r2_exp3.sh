#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lzb_decode_kernel -s 3 -c 1 -f -o $O/r2_k1_c2_full \
  python bench.py --config c2 --steps 2 --warmup 3 --no-verify --cpu-sample 64 > $O/r2_k1_c2_full.log 2>&1
( timeout 900 python tools/e2e_trace.py --config ns --steps 3 ) > $O/r2_e2e_trace_gated.txt 2>&1
( LZB_NO_GATE=1 timeout 900 python tools/e2e_trace.py --config ns --steps 2 ) > $O/r2_e2e_trace_nogate.txt 2>&1
( LZB_NO_MIRROR=1 timeout 900 python tools/e2e_trace.py --config ns --steps 2 ) > $O/r2_e2e_trace_nomirror.txt 2>&1
tail -3 $O/r2_k1_c2_full.log; grep -h "lzb_trace\|step" $O/r2_e2e_trace_*.txt
