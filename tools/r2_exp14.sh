#!/bin/bash
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2_exp14_pytest.txt
cat $O/r2_exp14_pytest.txt
