#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python tools/dbg_raw.py > $O/r2_exp11_dbg.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/dbg_raw.py > $O/r2_exp11_memcheck.txt 2>&1
cat $O/r2_exp11_dbg.txt; grep -v "^=========$" $O/r2_exp11_memcheck.txt | tail -40
