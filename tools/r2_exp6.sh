#!/bin/bash
# 8 GPUs: the north-star bench at N=8 (strong scaling + sharded runs) with the host-API timeline of every rank
O=gpurun_out
mkdir -p $O
( nvidia-smi topo -m; nproc; free -g; python -c "import os;print(os.cpu_count(), len(os.sched_getaffinity(0)))" ) > $O/r2_topo8.txt 2>&1
LZB_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2_exp6_ns_n8.json 2> $O/r2_exp6_ns_n8.err
head -30 $O/r2_topo8.txt; grep "lzb_trace" $O/r2_exp6_ns_n8.err | tail -40; grep -v "lzb_trace\|Warning\|UserWarning\|buf\[" $O/r2_exp6_ns_n8.err | tail -10; cat $O/r2_exp6_ns_n8.json
