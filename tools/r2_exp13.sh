#!/bin/bash
O=gpurun_out
timeout 300 python tools/dbg_raw.py > $O/r2_exp13_dbg.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "raw_decoders or cpp_host or config5 or golden or cuda_path" 2>&1 | tail -30 > $O/r2_exp13_pytest.txt
cat $O/r2_exp13_dbg.txt $O/r2_exp13_pytest.txt
