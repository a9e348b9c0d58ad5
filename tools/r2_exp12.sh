#!/bin/bash
O=gpurun_out
( LZB_FORCE_BIGLIT=1 timeout 300 python tools/dbg_big.py ) > $O/r2_exp12_big.txt 2>&1
timeout 600 compute-sanitizer --tool initcheck --print-limit 8 python tools/dbg_raw.py > $O/r2_exp12_initcheck.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 8 python tools/dbg_raw.py > $O/r2_exp12_racecheck.txt 2>&1
cat $O/r2_exp12_big.txt; grep -v "^=========$" $O/r2_exp12_initcheck.txt | head -60; grep -v "^=========$" $O/r2_exp12_racecheck.txt | head -40
