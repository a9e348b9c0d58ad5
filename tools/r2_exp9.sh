#!/bin/bash
O=gpurun_out
mkdir -p $O
V=build/variants
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > $O/r2_exp9_pytest.txt
timeout 600 python tools/kbench.py --config c5 --steps 5 $V/r2e.so $V/r2e_fillsc.so $V/r2f.so > $O/r2_exp9_c5.txt 2>&1
timeout 600 python tools/kbench.py --config c2 --steps 7 $V/r2c.so $V/r2f.so $V/r2c.so $V/r2f.so > $O/r2_exp9_c2.txt 2>&1
timeout 600 python tools/kbench.py --config c2 --streams 1024 --steps 7 $V/base.so $V/r2f.so > $O/r2_exp9_c2_1024.txt 2>&1
cat $O/r2_exp9_pytest.txt $O/r2_exp9_c5.txt $O/r2_exp9_c2.txt $O/r2_exp9_c2_1024.txt
