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

    # tensor form: one blob on the source, contiguous ranges per rank, slices sent / received in place
    import torch
    from lzma_rs_b200 import _native

    def decode_t(blob_t, in_off, out_off, out_t):
        m = len(in_off) - 1
        out_len, codes = np.zeros(m, dtype=np.uint64), np.zeros(m, dtype=np.int32)
        b, o = blob_t.numpy(), out_t.numpy()
        for i in range(m):
            r = oracle_py.lzma2_decompress(b[int(in_off[i]):int(in_off[i + 1])].tobytes())
            codes[i] = r.kind
            out_len[i] = len(r.out)
            o[int(out_off[i]):int(out_off[i]) + len(r.out)] = np.frombuffer(r.out, dtype=np.uint8)
        return out_len, codes

    if rank == 0:
        blob, in_off = _native.pack_streams(streams)
        res = sharding.decode_sharded_tensors(decode_t, torch.from_numpy(blob), in_off, [len(p) for p in plains], src=0)
        out_t, out_off, out_len, codes = res
        o = out_t.numpy()
        ok = ok and (codes == 0).all() and all(
            o[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes() == plains[i] for i in range(n))
    else:
        assert sharding.decode_sharded_tensors(decode_t, None, None, None, src=0) is None
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


def test_partition_contiguous_balances_and_covers():
    from lzma_rs_b200.sharding import partition_contiguous
    rng = np.random.default_rng(4)
    lens = rng.integers(0, 1 << 18, size=3000)
    off = np.zeros(3001, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    off += np.uint64(48)  # a batch that starts inside a larger buffer
    for world in (1, 2, 3, 8):
        r = partition_contiguous(off, world)
        assert r[0][0] == 0 and r[-1][1] == 3000 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        loads = np.array([int(off[h] - off[l]) for l, h in r])
        assert loads.max() - loads.min() <= 2 * lens.max()
    assert partition_contiguous(np.zeros(1, dtype=np.uint64), 4) == [(0, 0)] * 4


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


def _worker_small(rank, world, port, q):
    """Tensor form with more ranks than the data needs: some ranks get no stream at all, or only empty ones."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import corpus
    import oracle_py
    from lzma_rs_b200 import _native, sharding

    def decode_t(blob_t, in_off, out_off, out_t):
        m = len(in_off) - 1
        out_len, codes = np.zeros(m, dtype=np.uint64), np.zeros(m, dtype=np.int32)
        b, o = blob_t.numpy(), out_t.numpy()
        for i in range(m):
            r = oracle_py.lzma2_decompress(b[int(in_off[i]):int(in_off[i + 1])].tobytes())
            codes[i], out_len[i] = r.kind, len(r.out)
            o[int(out_off[i]):int(out_off[i]) + len(r.out)] = np.frombuffer(r.out, dtype=np.uint8)
        return out_len, codes

    ok = True
    for sizes in ([50_000, 10, 0], [0, 0], [7], []):
        if rank == 0:
            plains = [corpus.mixed_text(40 + i, s) for i, s in enumerate(sizes)]
            streams = [corpus.raw_lzma2(p) for p in plains]
            blob, in_off = _native.pack_streams(streams)
            res = sharding.decode_sharded_tensors(decode_t, torch.from_numpy(blob), in_off, [len(p) for p in plains], src=0)
            out_t, out_off, out_len, codes = res
            o = out_t.numpy()
            ok = ok and len(out_len) == len(plains) and bool((codes == 0).all()) and all(
                o[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes() == plains[i] for i in range(len(plains)))
        else:
            sharding.decode_sharded_tensors(decode_t, None, None, None, src=0)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_tensor_sharding_with_idle_ranks_world3():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_small, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
