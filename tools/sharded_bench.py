#!/usr/bin/env python3
"""tools/sharded_bench.py -- north_star's multi-GPU data path, timed: the whole batch sits in rank 0's HBM, compressed
shards are scattered over NCCL (NVLink), every rank decodes its shard, decoded shards are gathered back into rank 0's
output blob.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/sharded_bench.py [--config c2] [--streams-per-gpu 4096] [--steps 5]

Prints one JSON line on rank 0: whole-job decompressed GB/s including scatter and gather (device timers, max over
ranks), with the per-phase split of rank 0.  Output is compared with the plaintexts."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--streams-per-gpu", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from lzma_rs_b200 import Context, _native, sharding
    ctx = Context(local)
    fn = sharding.cuda_decode_fn(ctx, 1)
    bench.CFG = bench.CONFIGS[a.config]
    n = a.streams_per_gpu * world
    blob_t = in_off = caps = plain = None
    if rank == 0:
        distinct = min(n, 4096 if bench.CFG["stream_bytes"] <= 65536 else 1024)
        comp, plain = bench.build_corpus(0, n, distinct, min(32, os.cpu_count() or 1))
        blob, in_off = _native.pack_streams(comp)
        blob_t = torch.from_numpy(blob).cuda()
        caps = [len(p) for p in plain]
    dist.barrier()
    times = []
    res = None
    for step in range(a.steps + 2):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        res = sharding.decode_sharded_tensors(fn, blob_t, in_off, caps, src=0)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if step >= 2:
            times.append(float(dt[0]))
    if rank == 0:
        out_t, out_off, out_len, codes = res
        assert (codes == 0).all()
        host = out_t.cpu().numpy()
        for i in range(0, n, 61):
            o = int(out_off[i])
            assert host[o:o + int(out_len[i])].tobytes() == plain[i], i
        total = int(sum(caps))
        med = sorted(times)[len(times) // 2]
        print(json.dumps({"metric": "decompressed GB/s incl. NCCL scatter of compressed shards and gather of outputs",
                          "value": total / med / 1e9, "unit": "GB/s", "n_gpus": world, "ms_per_step": med * 1e3,
                          "streams": n, "compressed_bytes": int(in_off[-1]), "decompressed_bytes": total,
                          "config": bench.CFG["name"].format(n=a.streams_per_gpu), "verified": "sampled streams bit-exact",
                          "scaling": "weak", "timing": "wall clock around scatter+decode+gather, cuda-synchronised, max over ranks"}))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
