#!/bin/bash
# 2 GPUs: IPC / multi-device tests, then the north-star bench at N=2 (strong scaling + sharded runs)
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r2_topo2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -15 > $O/r2_exp5_pytest.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2_exp5_ns_n2.json 2> $O/r2_exp5_ns_n2.err
cat $O/r2_topo2.txt $O/r2_exp5_pytest.txt; tail -20 $O/r2_exp5_ns_n2.err; cat $O/r2_exp5_ns_n2.json
