"""N>1 host logic on CPU: world_size-2 gloo processes scatter compressed shards, decode locally and gather.
The per-rank decode is the oracle here (no GPU in the CPU tier); on the GPU box the same plumbing runs over NCCL
with the CUDA path (tests/test_gpu_parity.py::test_sharded_two_ranks when two GPUs are visible)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import corpus
    import oracle_py
    from lzma_rs_b200 import sharding
    n = 23
    sizes = [0, 1, 500, 70_000, 3000, 120_000] * 4
    plains = [corpus.mixed_text(700 + i, sizes[i]) for i in range(n)]
    streams = [corpus.raw_lzma2(p) for p in plains] if rank == 0 else None

    def decode(local):
        return [oracle_py.lzma2_decompress(s).out for s in local]

    out = sharding.decode_sharded(decode, streams, n, src=0)
    ok = True
    if rank == 0:
        ok = out == plains
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_lpt_partition_balances_and_covers():
    from lzma_rs_b200.sharding import lpt_partition
    rng = np.random.default_rng(3)
    lens = rng.integers(1, 1 << 20, size=1000)
    for world in (1, 2, 4, 8):
        parts = lpt_partition(lens, world)
        allidx = np.sort(np.concatenate(parts))
        assert (allidx == np.arange(1000)).all()
        loads = np.array([lens[p].sum() for p in parts])
        assert loads.max() - loads.min() <= lens.max()  # LPT bound
    assert [len(p) for p in lpt_partition([], 4)] == [0, 0, 0, 0]


def test_scatter_decode_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
