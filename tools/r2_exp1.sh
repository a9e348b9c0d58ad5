#!/bin/bash
# first GPU call of round 2: box facts + K1 variant A/B (tools/kbench.py) + GPU parity tests of the new default build
O=gpurun_out
mkdir -p $O
( nproc; free -g; nvidia-smi -L; nvidia-smi topo -m; cat /sys/fs/cgroup/cpu.max; python -c "import os;print(os.cpu_count(), len(os.sched_getaffinity(0)))"; ulimit -l ) > $O/r2_box.txt 2>&1
V=build/variants
timeout 900 python tools/kbench.py --config c2 --steps 7 $V/base.so $V/all.so $V/direct.so $V/trees.so $V/trees2.so $V/copy.so $V/checks.so $V/state.so $V/all_trees2.so $V/all_notrees.so $V/all_nocopy.so $V/all_nostate.so $V/base.so > $O/r2_exp1_c2.txt 2>&1
timeout 900 python tools/kbench.py --config ns --steps 3 --streams 8192 $V/base.so $V/all.so $V/all_trees2.so $V/all_notrees.so > $O/r2_exp1_ns.txt 2>&1
timeout 600 python tools/kbench.py --config c5 --steps 3 $V/base.so $V/all.so > $O/r2_exp1_c5.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r2_exp1_pytest.txt
cat $O/r2_box.txt $O/r2_exp1_c2.txt $O/r2_exp1_ns.txt $O/r2_exp1_c5.txt $O/r2_exp1_pytest.txt
