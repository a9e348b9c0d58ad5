#!/bin/bash
O=gpurun_out
mkdir -p $O
V=build/variants
timeout 900 python tools/kbench.py --config c2 --steps 7 $V/all.so $V/r2b.so $V/r2b_nostate.so $V/all.so $V/r2b.so > $O/r2_exp4_c2.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_exp4_pytest.txt
timeout 1500 python bench.py --steps 4 --warmup 3 > $O/r2_exp4_ns_full.json 2> $O/r2_exp4_ns_full.err
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 > $O/r2_exp4_c2.json 2> $O/r2_exp4_c2.err
( LZB_TRACE=1 timeout 600 python bench.py --config c3 --steps 3 --warmup 3 --cpu-sample 256 ) > $O/r2_exp4_c3.json 2> $O/r2_exp4_c3.err
cat $O/r2_exp4_c2.txt $O/r2_exp4_pytest.txt; tail -3 $O/r2_exp4_ns_full.err; cat $O/r2_exp4_ns_full.json; tail -3 $O/r2_exp4_c2.err; cat $O/r2_exp4_c2.json; grep lzb_trace $O/r2_exp4_c3.err | tail -3; cat $O/r2_exp4_c3.json
