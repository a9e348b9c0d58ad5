"""Multi-GPU host logic: independent streams sharded across ranks (one process per GPU).

Streams never talk to each other, so the data path needs no collective: each rank decodes its
own shard with its own `lzb_ctx`.  This module holds what north_star asks NCCL for: partitioning a batch that
lives on one rank, scattering the compressed shards and gathering the decoded outputs, over `torch.distributed`
(backend nccl on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def lpt_partition(lengths, world):
    """Longest-processing-time greedy: sort by compressed length (the best cheap proxy for decode time: time is
    proportional to decoded decisions), always give the next stream to the least loaded rank.
    Returns a list of `world` int64 index arrays."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    loads = np.zeros(world, dtype=np.int64)
    bins = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(loads))
        bins[r].append(int(i))
        loads[r] += int(lengths[i]) + 64  # a per-stream constant keeps empty streams spread as well
    return [np.asarray(sorted(b), dtype=np.int64) for b in bins]


def _dev(group=None):
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def _pack(items):
    lens = np.fromiter((len(x) for x in items), dtype=np.int64, count=len(items))
    blob = np.empty(int(lens.sum()), dtype=np.uint8)
    o = 0
    for x, n in zip(items, lens):
        blob[o:o + n] = np.frombuffer(x, dtype=np.uint8)
        o += int(n)
    return blob, lens


def _unpack(blob, lens):
    out, o = [], 0
    for n in lens:
        out.append(blob[o:o + int(n)].tobytes())
        o += int(n)
    return out


def scatter_streams(streams, src=0, group=None):
    """`streams` (list of bytes) is only read on `src`.  Returns (global indices of this rank's streams, streams)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _dev(group)
    if rank == src:
        parts = lpt_partition([len(s) for s in streams], world)
        meta = [(p.tolist(), [len(streams[i]) for i in p]) for p in parts]
    else:
        meta = None
    box = [None]
    dist.scatter_object_list(box, meta if rank == src else None, src=src, group=group)  # tiny: indices + lengths
    idx, lens = box[0]
    total = int(sum(lens))
    if rank == src:
        mine = None
        for r in range(world):
            blob, _ = _pack([streams[i] for i in parts[r]])
            if r == src:
                mine = blob
            elif len(blob):
                dist.send(torch.from_numpy(blob).to(dev), dst=r, group=group)
        blob = mine
    else:
        t = torch.empty(total, dtype=torch.uint8, device=dev)
        if total:
            dist.recv(t, src=src, group=group)
        blob = t.cpu().numpy()
    return np.asarray(idx, dtype=np.int64), _unpack(blob, lens)


def gather_outputs(indices, outputs, n_total, dst=0, group=None):
    """Inverse of scatter_streams: rank `dst` gets the list of all outputs in the original order (others: None)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _dev(group)
    lens = [len(o) for o in outputs]
    metas = [None] * world if rank == dst else None
    dist.gather_object((list(map(int, indices)), lens), metas, dst=dst, group=group)
    if rank != dst:
        blob, _ = _pack(outputs)
        if len(blob):
            dist.send(torch.from_numpy(blob).to(dev), dst=dst, group=group)
        return None
    result = [None] * n_total
    for r in range(world):
        idx, ls = metas[r]
        if r == dst:
            parts = outputs
        else:
            t = torch.empty(int(sum(ls)), dtype=torch.uint8, device=dev)
            if t.numel():
                dist.recv(t, src=r, group=group)
            parts = _unpack(t.cpu().numpy(), ls)
        for i, p in zip(idx, parts):
            result[i] = p
    return result


def decode_sharded(decode_fn, streams, n_total, src=0, group=None):
    """scatter -> local decode -> gather.  `decode_fn(list of bytes) -> list of bytes` is the per-rank decode
    (on a GPU box: lambda s: [r.data for r in ctx.decode_batch(fmt, s)])."""
    idx, mine = scatter_streams(streams, src, group)
    outs = decode_fn(mine) if len(mine) else []
    return gather_outputs(idx, outs, n_total, src, group)


# ------------------------------------------------------------------------------------------------
# Device-resident form: the batch is ONE blob tensor on the source rank (a GPU tensor under NCCL).  Ranks get
# CONTIGUOUS ranges of streams, balanced by compressed bytes, so that scatter and gather are plain slices of the
# input / output blobs -- no packing pass, one NCCL send/recv per peer and direction over NVLink.
# ------------------------------------------------------------------------------------------------
def partition_contiguous(in_off, world):
    """Stream index ranges [lo, hi) per rank: cut points where the cumulative compressed size crosses k/world."""
    in_off = np.asarray(in_off, dtype=np.uint64)
    n = len(in_off) - 1
    base, total = int(in_off[0]), int(in_off[-1] - in_off[0])
    cuts = [0]
    for k in range(1, world):
        target = base + total * k // world
        c = int(np.searchsorted(in_off, np.uint64(target), side="left"))
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def align16(x):
    return (np.asarray(x, dtype=np.uint64) + np.uint64(15)) // np.uint64(16) * np.uint64(16)


def decode_sharded_tensors(decode_fn, blob, in_off, caps, src=0, group=None):
    """scatter -> local decode -> gather on tensors.

    On `src`: `blob` = uint8 tensor holding all streams (on the device under NCCL), `in_off` = n+1 uint64 offsets
    into it, `caps` = n output capacities.  Other ranks pass None.  `decode_fn(blob_t, in_off, out_off, out_t)`
    decodes this rank's shard in place (blob_t / out_t live where the backend wants them; offsets are numpy
    uint64 arrays relative to those tensors) and returns (out_len, codes) numpy arrays.
    Returns on `src`: (out tensor, out_off, out_len, codes) for the whole batch; elsewhere None.
    All metadata travels as int64 tensors (two small broadcasts, one small send per peer): no pickling on the path."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _dev(group)
    head = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == src:
        in_off = np.asarray(in_off, dtype=np.uint64)
        n = len(in_off) - 1
        out_off_all = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(align16(caps), out=out_off_all[1:])
        ranges = partition_contiguous(in_off, world)
        cuts = np.array([r[0] for r in ranges] + [n], dtype=np.int64)
        head[0] = n
        meta = torch.from_numpy(np.concatenate([cuts, in_off.astype(np.int64), out_off_all.astype(np.int64)])).to(dev)
    dist.broadcast(head, src=src, group=group)
    n = int(head[0])
    if rank != src:
        meta = torch.empty(world + 1 + 2 * (n + 1), dtype=torch.int64, device=dev)
    dist.broadcast(meta, src=src, group=group)
    m = meta.cpu().numpy()
    cuts, in_all, out_all_off = m[:world + 1], m[world + 1:world + 1 + n + 1].astype(np.uint64), m[world + 2 + n:].astype(np.uint64)
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    my_in, my_out = in_all[lo:hi + 1], out_all_off[lo:hi + 1]
    n_loc = hi - lo
    in_lo, in_hi, out_lo, out_hi = int(my_in[0]), int(my_in[-1]), int(my_out[0]), int(my_out[-1])
    # 16 bytes of slack behind every shard: the kernels read whole aligned words
    if rank == src:
        out_t = torch.zeros(int(out_all_off[-1]) + 16, dtype=torch.uint8, device=dev)
        # one batched group of sends: all peers are fed concurrently (unbatched NCCL P2P calls are serialised)
        ops = []
        for r in range(world):
            a, b = int(in_all[cuts[r]]), int(in_all[cuts[r + 1]])
            if r != src and b > a:
                ops.append(dist.P2POp(dist.isend, blob[a:b], r, group))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        # the source decodes straight out of / into the full blobs.  The kernels read whole aligned words: when the
        # source's own range reaches the end of `blob` without 16 bytes of slack behind it, decode a padded copy instead
        shard, shard_base, out_base = blob, 0, 0
        if in_hi > in_lo and in_hi + 16 > blob.numel():
            shard = torch.zeros(in_hi - in_lo + 16, dtype=torch.uint8, device=dev)
            shard[:in_hi - in_lo] = blob[in_lo:in_hi]
            shard_base = in_lo
    else:
        shard = torch.zeros(in_hi - in_lo + 16, dtype=torch.uint8, device=dev)
        if in_hi > in_lo:
            dist.recv(shard[:in_hi - in_lo], src=src, group=group)
        shard_base, out_base = in_lo, out_lo
        out_t = torch.zeros(out_hi - out_lo + 16, dtype=torch.uint8, device=dev)
        reqs = []
    if n_loc:
        out_len, codes = decode_fn(shard, my_in - np.uint64(shard_base), my_out - np.uint64(out_base), out_t)
    else:
        out_len, codes = np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.int32)
    for q in reqs:
        q.wait()
    # gather: decoded shards are contiguous slices of the full output blob; per-stream results ride along as int64
    if rank != src:
        res = torch.from_numpy(np.concatenate([np.asarray(out_len, dtype=np.int64), np.asarray(codes, dtype=np.int64)])).to(dev)
        ops = []
        if n_loc:
            ops.append(dist.P2POp(dist.isend, res, src, group))
        if out_hi > out_lo:
            ops.append(dist.P2POp(dist.isend, out_t[:out_hi - out_lo], src, group))
        for q in (dist.batch_isend_irecv(ops) if ops else []):
            q.wait()
        return None
    all_len, all_codes = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.int32)
    all_len[lo:hi], all_codes[lo:hi] = out_len, codes
    # one batched group of receives, posted AFTER the source's own decode: all peers deliver concurrently, and the NCCL
    # kernel does not sit on SMs the persistent decode kernel needs (it owns every register of an SM it runs on)
    ops, metas = [], []
    for r in range(world):
        a, b = int(cuts[r]), int(cuts[r + 1])
        if r == src or b == a:
            continue
        res = torch.empty(2 * (b - a), dtype=torch.int64, device=dev)
        ops.append(dist.P2POp(dist.irecv, res, r, group))
        oa, ob = int(out_all_off[a]), int(out_all_off[b])
        if ob > oa:
            ops.append(dist.P2POp(dist.irecv, out_t[oa:ob], r, group))
        metas.append((a, b, res))
    for q in (dist.batch_isend_irecv(ops) if ops else []):
        q.wait()
    for a, b, res in metas:
        rn = res.cpu().numpy()
        all_len[a:b], all_codes[a:b] = rn[:b - a].astype(np.uint64), rn[b - a:].astype(np.int32)
    return out_t, out_all_off, all_len, all_codes


def cuda_decode_fn(ctx, fmt=1, options=None):
    """decode_fn for decode_sharded_tensors on a GPU box: lzb_decode_batch_device on this rank's context."""
    import ctypes as C
    from . import _native, decompress

    def fn(blob_t, in_off, out_off, out_t):
        n = len(in_off) - 1
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        out_len, consumed = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        st = np.zeros(n, dtype=_native.STATUS_DTYPE)
        opt = (options or decompress.Options())._native()
        # the shard was received by NCCL in order of torch's current stream; the library works on its own stream
        torch.cuda.current_stream().synchronize()
        rc = _native.load().lzb_decode_batch_device(ctx.handle, fmt, C.byref(opt), blob_t.data_ptr(), in_off.ctypes.data, n,
                                                    out_t.data_ptr(), out_off.ctypes.data, out_len.ctypes.data,
                                                    consumed.ctypes.data, st.ctypes.data, None)
        if rc != 0:
            raise RuntimeError(f"lzb_decode_batch_device failed rc={rc}: {ctx.last_error()}")
        return out_len, st["code"].astype(np.int32)

    return fn
