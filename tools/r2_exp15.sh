#!/bin/bash
# 4 GPUs: the 2-GPU tests, then the north-star bench at N=2 and N=4
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharded.py -q 2>&1 | tail -15 > $O/r2_exp15_pytest.txt
LZB_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 5 --warmup 3 > $O/r2_exp15_ns_n4.json 2> $O/r2_exp15_ns_n4.err
CUDA_VISIBLE_DEVICES=0,1 LZB_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2_exp15_ns_n2.json 2> $O/r2_exp15_ns_n2.err
cat $O/r2_exp15_pytest.txt; grep -v "lzb_trace\|Warning\|UserWarning\|buf\[\|\*\*\*\|OMP_NUM" $O/r2_exp15_ns_n4.err | tail -5; grep -o '{"metric.*' $O/r2_exp15_ns_n4.json | cut -c1-300; grep -o '{"metric.*' $O/r2_exp15_ns_n2.json | cut -c1-300
