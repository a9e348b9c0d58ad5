#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_abi2.py -x -q 2>&1 | tail -30 > $O/r2_exp2_pytest.txt
timeout 900 python bench.py --streams 8192 --distinct 512 --steps 3 --warmup 3 > $O/r2_exp2_ns_small.json 2> $O/r2_exp2_ns_small.err
timeout 1500 python bench.py --steps 5 --warmup 3 > $O/r2_exp2_ns_full.json 2> $O/r2_exp2_ns_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_exp2_ns_ref.json 2> $O/r2_exp2_ns_ref.err
cat $O/r2_exp2_pytest.txt; tail -5 $O/r2_exp2_ns_small.err; cat $O/r2_exp2_ns_small.json; tail -5 $O/r2_exp2_ns_full.err; cat $O/r2_exp2_ns_full.json; cat $O/r2_exp2_ns_ref.json
