#!/bin/bash
O=gpurun_out
mkdir -p $O
V=build/variants
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "raw_decoders or cpp_host" 2>&1 | tail -60 > $O/r2_exp10_pytest.txt
timeout 600 python tools/kbench.py --config c5 --steps 5 $V/r2e.so $V/r2e_fillsc.so $V/r2f.so > $O/r2_exp10_c5.txt 2>&1
timeout 600 python tools/kbench.py --config c2 --steps 7 $V/r2c.so $V/r2f.so $V/r2c.so $V/r2f.so > $O/r2_exp10_c2.txt 2>&1
cat $O/r2_exp10_pytest.txt $O/r2_exp10_c5.txt $O/r2_exp10_c2.txt
