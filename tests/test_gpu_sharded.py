"""GPU tier, N>1: one process per GPU; rank 0 scatters compressed shards over NCCL, every rank decodes its shard on
its own GPU through the C ABI, outputs are gathered on rank 0 and checked bit-exact.  Skipped with < 2 GPUs."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import corpus
    from lzma_rs_b200 import Context, sharding
    ctx = Context(rank)
    n = 97
    plains = [corpus.mixed_text(5000 + i, [0, 1, 700, 65536, 200_000, 3000][i % 6]) for i in range(n)]
    streams = [corpus.raw_lzma2(p, dict_size=1 << 20) for p in plains] if rank == 0 else None

    def decode(local):
        res = ctx.decode_batch(1, local)
        assert all(r.ok for r in res)
        return [r.data for r in res]

    out = sharding.decode_sharded(decode, streams, n, src=0)
    ok = (out == plains) if rank == 0 else True
    # tensor form: the batch is one blob in rank 0's HBM; contiguous slices travel over NCCL, outputs come back in place
    import numpy as np
    from lzma_rs_b200 import _native
    fn = sharding.cuda_decode_fn(ctx, 1)
    if rank == 0:
        blob, in_off = _native.pack_streams(streams)
        res = sharding.decode_sharded_tensors(fn, torch.from_numpy(blob).cuda(), in_off, [len(p) for p in plains], src=0)
        out_t, out_off, out_len, codes = res
        o = out_t.cpu().numpy()
        ok = ok and bool((codes == 0).all()) and all(
            o[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes() == plains[i] for i in range(n))
    else:
        sharding.decode_sharded_tensors(fn, None, None, None, src=0)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_sharded_two_ranks_nccl():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
