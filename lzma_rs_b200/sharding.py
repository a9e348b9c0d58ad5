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
