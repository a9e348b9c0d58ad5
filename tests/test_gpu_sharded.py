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


def _worker_p2p(rank, world, port, q):
    """The batch lives in rank 0's HBM; the other ranks map its blobs (CUDA IPC) and run lzb_decode_batch_peer on their
    stream range: input pulled over NVLink behind the gate, output pages stored into rank 0's blob by K1."""
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import corpus
    from lzma_rs_b200 import Context, _native, sharding
    lib = _native.load()
    ctx = Context(rank)
    rng = np.random.default_rng(5)
    base_plain = [corpus.mixed_text(7000 + i, int(rng.integers(30_000, 200_000))) for i in range(48)]
    base = [corpus.raw_lzma2(p, dict_size=1 << 20) for p in base_plain]
    n = 900
    streams = [base[i % 48] for i in range(n)]
    plains = [base_plain[i % 48] for i in range(n)]
    streams[450] = streams[450][:200]  # one stream that ends in an error, in rank 1's half
    blob, in_off = _native.pack_streams(streams)
    assert int(in_off[-1]) > 40 << 20  # both halves arm the input gate
    caps = np.array([len(p) for p in plains], dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((caps + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
    handles = [None, None]
    if rank == 0:
        full_in = torch.from_numpy(blob).cuda()
        full_out = torch.zeros(int(out_off[-1]) + 16, dtype=torch.uint8, device="cuda")
        hs = []
        for t in (full_in, full_out):
            h = _native.IpcHandle()
            assert lib.lzb_ipc_export(ctx.handle, t.data_ptr(), t.numel(), C.byref(h)) == 0, ctx.last_error()
            hs.append(bytes(h))
        handles = hs
    dist.broadcast_object_list(handles, src=0)
    if rank == 0:
        p_in, p_out = full_in.data_ptr(), full_out.data_ptr()
    else:
        ptrs = []
        for raw in handles:
            p = C.c_void_p()
            h = _native.IpcHandle.from_buffer_copy(raw)
            assert lib.lzb_ipc_open(ctx.handle, C.byref(h), C.byref(p)) == 0, ctx.last_error()
            ptrs.append(p.value)
        p_in, p_out = ptrs
    lo, hi = sharding.partition_contiguous(in_off, world)[rank]
    m = hi - lo
    sub_in, sub_out = np.ascontiguousarray(in_off[lo:hi + 1]), np.ascontiguousarray(out_off[lo:hi + 1])
    out_len, cons = np.zeros(m, dtype=np.uint64), np.zeros(m, dtype=np.uint64)
    st = np.zeros(m, dtype=_native.STATUS_DTYPE)
    opt = _native.make_options()
    ok = True
    for _ in range(2):
        if rank == 0:
            full_out.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            rc = lib.lzb_decode_batch_device(ctx.handle, 1, C.byref(opt), p_in, sub_in.ctypes.data, m, p_out,
                                             sub_out.ctypes.data, out_len.ctypes.data, cons.ctypes.data, st.ctypes.data, None)
        else:
            rc = lib.lzb_decode_batch_peer(ctx.handle, 1, C.byref(opt), p_in, sub_in.ctypes.data, m, p_out,
                                           sub_out.ctypes.data, out_len.ctypes.data, cons.ctypes.data, st.ctypes.data)
        ok = ok and rc == 0
        torch.cuda.synchronize()
        dist.barrier()
        codes = torch.from_numpy(st["code"].astype(np.int64)).cuda()
        lens = torch.from_numpy(out_len.astype(np.int64)).cuda()
        if rank == 0:
            all_codes, all_lens = [codes], [lens]
            for r in range(1, world):
                a, b = sharding.partition_contiguous(in_off, world)[r]
                c, l = torch.empty(b - a, dtype=torch.int64, device="cuda"), torch.empty(b - a, dtype=torch.int64, device="cuda")
                dist.recv(c, src=r)
                dist.recv(l, src=r)
                all_codes.append(c)
                all_lens.append(l)
            codes_h, lens_h = torch.cat(all_codes).cpu().numpy(), torch.cat(all_lens).cpu().numpy()
            host = full_out.cpu().numpy()
            import oracle_py as oracle
            bad = oracle.lzma2_decompress(streams[450])
            for i in range(n):
                got = host[int(out_off[i]):int(out_off[i]) + int(lens_h[i])].tobytes()
                if i == 450:
                    ok = ok and codes_h[i] != 0 and got == bad.out
                else:
                    ok = ok and codes_h[i] == 0 and got == plains[i]
        else:
            dist.send(codes, dst=0)
            dist.send(lens, dst=0)
    if rank != 0:
        assert lib.lzb_ipc_close(ctx.handle, p_in) == 0 and lib.lzb_ipc_close(ctx.handle, p_out) == 0
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def _run_two(worker):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_peer_decode_two_ranks_ipc():
    _run_two(_worker_p2p)


def test_multi_device_two_gpus():
    """lzb_decode_batch_multi over two real devices in one process: every device uploads its own stream range."""
    import ctypes as C
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import corpus
    from lzma_rs_b200 import _native
    lib = _native.load()
    rng = np.random.default_rng(6)
    base_plain = [corpus.mixed_text(7100 + i, int(rng.integers(30_000, 200_000))) for i in range(48)]
    base = [corpus.raw_lzma2(p, dict_size=1 << 20) for p in base_plain]
    n = 900
    streams, plains = [base[i % 48] for i in range(n)], [base_plain[i % 48] for i in range(n)]
    blob, in_off = _native.pack_streams(streams)
    caps = np.array([len(p) for p in plains], dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((caps + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
    h_in = torch.from_numpy(blob).pin_memory()
    h_out = torch.zeros(int(out_off[-1]) + 16, dtype=torch.uint8).pin_memory()
    m = C.c_void_p()
    assert lib.lzb_create_multi(C.byref(m), None, 0) == 0
    assert lib.lzb_multi_device_count(m) == torch.cuda.device_count()
    out_len, cons = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=_native.STATUS_DTYPE)
    opt = _native.make_options()
    rc = lib.lzb_decode_batch_multi(m, 1, C.byref(opt), h_in.data_ptr(), in_off.ctypes.data, n, h_out.data_ptr(),
                                    out_off.ctypes.data, out_len.ctypes.data, cons.ctypes.data, st.ctypes.data, None)
    assert rc == 0, lib.lzb_multi_last_error(m)
    hv = h_out.numpy()
    assert (st["code"] == 0).all()
    for i in range(n):
        assert hv[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes() == plains[i], i
    lib.lzb_destroy_multi(m)


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
