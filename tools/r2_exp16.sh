#!/bin/bash
# 8 GPUs: the north-star bench at N=8 and BASELINE config 3 at its stated scale (65 536 x 256 KiB over 8 GPUs)
O=gpurun_out
mkdir -p $O
LZB_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2_exp16_ns_n8.json 2> $O/r2_exp16_ns_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --config c3 --steps 5 --warmup 3 > $O/r2_exp16_c3_n8.json 2> $O/r2_exp16_c3_n8.err
grep "lzb_trace n=" $O/r2_exp16_ns_n8.err | tail -8; grep -o '{"metric.*' $O/r2_exp16_ns_n8.json | cut -c1-250; grep -v "Warning\|UserWarning\|\*\*\*\|OMP_NUM\|^$" $O/r2_exp16_c3_n8.err | tail -4; grep -o '{"metric.*' $O/r2_exp16_c3_n8.json | cut -c1-400
