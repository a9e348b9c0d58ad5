#!/usr/bin/env python3
"""bench.py -- decompressed GB/s of the many-stream LZMA2 decode path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config ns|c2|c3|c4|c5|c6]

Default workload = the batch BASELINE.json's north_star quotes the target on: 65 536 independent raw-LZMA2 streams
(lc3 lp0 pb2, dict 1 MiB, liblzma preset 6, seeded mixed literal/match text), decompressed sizes log-uniform in
[64 KiB, 1 MiB] -- about 11.6 GB in, 23.2 GB out -- tiled from 4 096 distinct streams.  The batch is FIXED and
STRONG-scaled: under torchrun (one process per GPU) rank r decodes a contiguous stream range holding 1/N of the
compressed bytes (streams are independent: SURVEY.md 8(e), no data-path collective).  `--config c2` etc. are the other
BASELINE.json configurations (per-GPU batches, weak scaling, informational).

A step = one pass of the hot path over the rank's shard.  Every timed step starts from a zeroed output buffer and is
compared with the plaintexts on the device before the next one starts (outside the timed intervals).
  value   : kernel-only, shards resident in HBM; CUDA events around every step on the launching stream, summed, max
            over ranks.
  e2e     : the same shards through the reference-facing C-ABI call lzb_decode_batch with pinned HOST buffers -- every
            rank uploads its own shard over its own PCIe link, K1 streams finished output pages back (H2D + scan +
            decode + D2H inside the timed region).
  sharded : (N > 1) the whole batch starts and ends in rank 0's HBM.  `p2p_fused`: the other ranks map rank 0's blobs
            (CUDA IPC over NVLink) and run lzb_decode_batch_peer: compressed bytes pulled in chunks behind K1's input
            gate, output pages stored into rank 0's blob by the decode kernel -- scatter + decode + gather in one launch.
            `nccl_scatter_gather`: the plain torch.distributed send/recv form (lzma_rs_b200/sharding.py), unoverlapped.
  roofline     : algorithmic bytes (compressed read once + decompressed written once) / kernel time vs the measured
                 HBM copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline : the C oracle (line-by-line restatement of lzma-rs's src/decode, oracle/) on the host cores, plus
                 liblzma (the reference's own differential oracle, tests/lzma.rs:109-114) on the same streams and threads.
--impl reference times that CPU path alone (the Rust reference cannot be built here: no rustc/cargo in the image).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

L2_BYTES = 126 * 1000 * 1000

# `ns` (default) = the north-star batch, strong-scaled (bench_ns below).  The others are BASELINE.json's configs as
# per-GPU batches (weak scaling, informational): c2 was the round-1 benchmark line.
CONFIGS = {
    "c2": dict(index=2, stream_bytes=65536, dict_size=1 << 18, streams=4096, kind="mixed",
               name="C2: {n} independent raw LZMA2 streams x 65536 B (lc3 lp0 pb2, dict 256 KiB, seeded mixed "
                    "literal/match text, liblzma preset 6) per GPU"),
    "c3": dict(index=3, stream_bytes=262144, dict_size=1 << 20, streams=8192, kind="mixed",
               name="C3 shard: {n} independent raw LZMA2 streams x 262144 B (dict 1 MiB, 3 chunks each) per GPU"),
    "c5": dict(index=5, stream_bytes=262144, dict_size=1 << 20, streams=8192, kind="rep0",
               name="C5: {n} raw LZMA2 streams x 262144 B of all-overlapping rep0 matches (dist=1, len=273) per GPU"),
    "ns": dict(index=6, stream_bytes=0, dict_size=1 << 20, streams=65536, kind="mixed",
               name="NS: {n} independent raw LZMA2 streams (lc3 lp0 pb2, dict 1 MiB, liblzma preset 6, seeded mixed "
                    "literal/match text), decompressed sizes log-uniform in [64 KiB, 1 MiB], tiled from {d} distinct "
                    "streams; ONE fixed batch, strong-scaled over the GPUs"),
    "c6": dict(index=7, stream_bytes=1 << 20, dict_size=1 << 20, streams=2048, kind="stored",
               name="C6: {n} raw LZMA2 streams x 1 MiB of stored chunks only (16 x `01 FF FF` + 64 KiB, the only LZMA2 the "
                    "reference's own encoder writes, src/encode/lzma2.rs:4-26) per GPU -- the byte-bound end of the path"),
    "c4": dict(index=4, stream_bytes=1 << 20, dict_size=1 << 20, streams=1024, kind="xz",
               name="C4: {n} .xz files x 1 MiB (4 blocks of 256 KiB each, LZMA2 filter, CRC32 block check) per GPU, "
                    "host API only (container walk on the host, K1 decode + K3 CRC on the GPU)"),
}
CFG = CONFIGS["ns"]


def _one_stream(args):
    import corpus
    seed, size, dict_size, kind = args
    if size == 0:  # north-star sweep: size log-uniform in [64 KiB, 1 MiB], a function of the seed
        size = int(65536 * 16 ** np.random.default_rng(seed ^ 0x5EED).random())
    if kind == "xz":
        plain = corpus.mixed_text(seed, size)
        return corpus.xz_file(plain, block_size=1 << 18, check=corpus.CHECK_CRC32, dict_size=dict_size), plain
    if kind == "stored":
        plain = np.random.default_rng(seed).bytes(size)
        return corpus.stored_lzma2(plain), plain
    if kind == "rep0":
        plain = bytes([seed & 0xFF]) * size
        return corpus.rep0_stress_lzma2(size, byte=seed & 0xFF), plain
    plain = corpus.mixed_text(seed, size)
    return corpus.raw_lzma2(plain, dict_size=dict_size, preset=6), plain


def build_corpus(rank, n_streams, distinct, workers):
    """Seeds: config_index * 1_000_003 + stream_index (+ rank offset so ranks hold different data)."""
    import multiprocessing as mp
    distinct = min(distinct, n_streams)
    jobs = [(CFG["index"] * 1_000_003 + rank * 100_003 + i, CFG["stream_bytes"], CFG["dict_size"], CFG["kind"])
            for i in range(distinct)]
    if workers > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            base = pool.map(_one_stream, jobs, chunksize=max(1, distinct // (workers * 4)))
    else:
        base = [_one_stream(j) for j in jobs]
    comp = [base[i % distinct][0] for i in range(n_streams)]
    plain = [base[i % distinct][1] for i in range(n_streams)]
    return comp, plain


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc, self.thr = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
        self.thr.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thr.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def effective_cpus():
    """Host threads this process can really use: min(os.cpu_count(), sched affinity, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    return n, quota


def cpu_reference_run(comp, plain, steps, warmup, threads, sample_streams):
    """The lzma-rs-equivalent CPU path (C oracle) over a bounded sample of the workload, one stream per task.
    `threads` = candidate thread counts: the fastest is reported (a container CPU quota can make fewer threads
    than os.cpu_count() faster)."""
    import oracle_py
    from lzma_rs_b200 import _native
    k = min(sample_streams, len(comp))
    blob, in_off = _native.pack_streams(comp[:k])
    sizes = np.array([len(p) for p in plain[:k]], dtype=np.uint64)
    out_off = np.zeros(k + 1, dtype=np.uint64)
    np.cumsum(sizes, out=out_off[1:])
    total = int(out_off[-1])
    best = None
    for th in threads:
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out, out_len, kinds, failed = oracle_py.decompress_batch(1, blob, in_off, out_off, th)
            dt = time.perf_counter() - t0
            assert failed == 0
            if it >= warmup:
                times.append(dt)
        dt = float(np.mean(times))
        if best is None or dt < best[0]:
            best = (dt, th)
    assert out[int(out_off[0]):int(out_off[1])].tobytes() == plain[0]
    dt, th = best
    return total / dt / 1e9, dt, k, total, th


def effective_cores(threads_used, ncap, quota):
    """Cores the CPU baseline really had: its threads, capped by the affinity mask and the cgroup CPU quota."""
    c = min(threads_used, ncap)
    if quota:
        c = min(c, max(1, int(round(quota))))
    return c


def cpu_thread_candidates(ncpu, quota):
    c = {ncpu, max(1, ncpu // 2)}
    if quota:
        c |= {max(1, int(round(quota))), max(1, int(round(quota * 2)))}
    c |= {x for x in (16, 32, 64) if x <= ncpu}
    return sorted(c)


def bench_xz(a, lib, ctx, comp, plain, workload, world, rank):
    """Informational: BASELINE config 4 through the reference-facing host API (there is no device-resident XZ entry)."""
    import torch
    from lzma_rs_b200 import _native
    n = len(comp)
    blob, in_off = _native.pack_streams(comp)
    sizes = np.array([len(p) for p in plain], dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((sizes + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
    h_in = torch.from_numpy(blob).pin_memory()
    h_out = torch.empty(int(out_off[-1]) + 16, dtype=torch.uint8).pin_memory()
    out_len = np.zeros(n, dtype=np.uint64)
    cons = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=_native.STATUS_DTYPE)
    opt = _native.make_options()

    def step():
        r = lib.lzb_decode_batch(ctx.handle, _native.FMT_XZ, C.byref(opt), h_in.data_ptr(), in_off.ctypes.data, n,
                                 h_out.data_ptr(), out_off.ctypes.data, out_len.ctypes.data, cons.ctypes.data,
                                 st.ctypes.data)
        assert r == 0, (r, ctx.last_error())

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    assert (st["code"] == 0).all()
    hv = h_out.numpy()
    for i in range(0, n, 37):
        o = int(out_off[i])
        assert hv[o:o + int(out_len[i])].tobytes() == plain[i]
    total = int(sizes.sum())
    print(json.dumps({"metric": "decompressed GB/s (.xz files, host API, informational)", "value": total / dt / 1e9,
                      "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
                      "higher_is_better": True, "data": "synthetic", "dtype": "u8/u16/u32 integer",
                      "config": {"workload": workload, "compressed_bytes": int(in_off[-1]), "decompressed_bytes": total},
                      "e2e": {"value": total / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(in_off[-1]),
                              "d2h_bytes_per_step": total}}))
    ctx.close()


# ------------------------------------------------------------------------------------------------------------------
# north-star batch (default): one fixed batch of 65 536 streams, strong-scaled
# ------------------------------------------------------------------------------------------------------------------
NS_DISTINCT = 4096


def ns_config(total_streams, distinct):
    """The `config` object of the benchmark line -- identical in both arms (`--impl ours` / `--impl reference`)."""
    return {"workload": CONFIGS["ns"]["name"].format(n=total_streams, d=distinct), "streams_total": total_streams,
            "distinct_streams": distinct}


def ns_seed(j):
    return CONFIGS["ns"]["index"] * 1_000_003 + j


def ns_plain_len(j):
    """Decompressed size of distinct stream j: log-uniform in [64 KiB, 1 MiB], a function of the seed (_one_stream)."""
    return int(65536 * 16 ** np.random.default_rng(ns_seed(j) ^ 0x5EED).random())


def ns_generate(j_lo, j_hi, workers):
    """Distinct streams [j_lo, j_hi): (compressed, plaintext) pairs."""
    import multiprocessing as mp
    jobs = [(ns_seed(j), 0, CONFIGS["ns"]["dict_size"], "mixed") for j in range(j_lo, j_hi)]
    if workers > 1 and len(jobs) > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            return pool.map(_one_stream, jobs, chunksize=max(1, len(jobs) // (workers * 8)))
    return [_one_stream(j) for j in jobs]


def align16(x):
    return (np.asarray(x, dtype=np.int64) + 15) // 16 * 16


def ns_layout(comp_lens, plain_lens, total_streams, world):
    """Offsets of the tiled batch (stream i = distinct[i % D]) and the contiguous stream range of every rank (equal
    shares of the compressed bytes, sharding.partition_contiguous)."""
    from lzma_rs_b200 import sharding
    d = len(comp_lens)
    reps = -(-total_streams // d)
    in_off = np.zeros(total_streams + 1, dtype=np.uint64)
    np.cumsum(np.tile(np.asarray(comp_lens, dtype=np.uint64), reps)[:total_streams], out=in_off[1:])
    out_off = np.zeros(total_streams + 1, dtype=np.uint64)
    np.cumsum(np.tile(align16(plain_lens).astype(np.uint64), reps)[:total_streams], out=out_off[1:])
    return in_off, out_off, sharding.partition_contiguous(in_off, world)


def tile_slice(torch, base, a, b, pad=32):
    """Bytes [a, b) of the infinite repetition of the 1-D uint8 tensor `base`, as a new tensor (+ `pad` zero bytes)."""
    period = base.numel()
    kw = {"pin_memory": True} if (pad < 0) else {}
    res = torch.empty(b - a + abs(pad), dtype=torch.uint8, device=base.device, **kw)
    res[b - a:] = 0
    pos = a
    while pos < b:
        o = pos % period
        n = min(period - o, b - pos)
        res[pos - a:pos - a + n] = base[o:o + n]
        pos += n
    return res


def tile_equal(torch, got, base, a, b):
    """got[0 : b-a] == bytes [a, b) of the infinite repetition of `base`?"""
    period = base.numel()
    pos = a
    while pos < b:
        o = pos % period
        n = min(period - o, b - pos)
        if not torch.equal(got[pos - a:pos - a + n], base[o:o + n]):
            return False
        pos += n
    return True


def cpu_liblzma_run(comp, plain, threads, dict_size):
    """liblzma (Python's lzma module; the reference's own differential oracle, tests/lzma.rs:109-114) on the same streams
    with the same number of threads: one stream per task, the GIL is released inside the decompressor."""
    import lzma
    from concurrent.futures import ThreadPoolExecutor
    filt = [{"id": lzma.FILTER_LZMA2, "dict_size": dict_size}]

    def one(c):
        return len(lzma.LZMADecompressor(format=lzma.FORMAT_RAW, filters=filt).decompress(c))
    total = sum(len(p) for p in plain)
    best = None
    with ThreadPoolExecutor(threads) as ex:
        for it in range(3):
            t0 = time.perf_counter()
            got = sum(ex.map(one, comp))
            dt = time.perf_counter() - t0
            assert got == total
            if it and (best is None or dt < best):
                best = dt
    return total / best / 1e9


def cpu_baselines(comp, plain, sample, n_total, dict_size, steps=2, warmup=1):
    """cpu_baseline (C oracle, kind "port") + liblzma on a bounded sample; GB/s in total and per core."""
    ncap, quota = effective_cpus()
    cands = cpu_thread_candidates(ncap, quota)
    gbs, dt, k, total, used = cpu_reference_run(comp, plain, steps, warmup, cands, sample)
    cores = effective_cores(used, ncap, quota)
    try:
        hw_threads = len(os.sched_getaffinity(0)), os.cpu_count()
    except Exception:
        hw_threads = (None, os.cpu_count())
    obj = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "per_core_gbs": gbs / cores,
           "sample": f"{k} of {n_total} streams ({total} B out), one stream per task, best of thread counts {cands} -> "
                     f"{used} pthreads (affinity {hw_threads[0]}, os.cpu_count {hw_threads[1]}, cgroup quota {quota}); C "
                     "restatement of lzma-rs src/decode (oracle/; the reference is Rust, no rustc in the image)"}
    try:
        lz = cpu_liblzma_run(comp[:k], plain[:k], used, dict_size)
        obj2 = {"value": lz, "unit": "GB/s", "cores": cores, "per_core_gbs": lz / cores, "kind": "liblzma (xz-utils via "
                "Python's lzma module), NOT the reference: the decoder the reference's own tests diff against",
                "sample": f"same {k} streams, {used} threads"}
    except Exception as e:  # pragma: no cover
        obj2 = {"value": None, "error": repr(e)}
    return obj, obj2


def bench_ns(a, rank, local_rank, world):
    total_streams = a.streams
    distinct = min(a.distinct or NS_DISTINCT, total_streams)
    assert distinct % world == 0, "distinct streams must divide over the ranks"
    ncpu = os.cpu_count() or 1
    try:
        ncpu = min(ncpu, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    workers = max(1, min(32, ncpu // max(1, world)))
    per = distinct // world
    t_gen = time.perf_counter()
    part = ns_generate(rank * per, (rank + 1) * per, workers)  # before CUDA init (fork pool)
    t_gen = time.perf_counter() - t_gen

    import torch
    import torch.distributed as dist
    from lzma_rs_b200 import Context, _native, sharding
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.load()
    ctx = Context(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- the distinct streams of every rank, exchanged over NCCL: compressed (tight) and plaintext (16-byte padded)
    comp_lens_part = np.array([len(c) for c, _ in part], dtype=np.int64)
    plain_lens = np.array([ns_plain_len(j) for j in range(distinct)], dtype=np.int64)
    assert [len(p) for _, p in part] == plain_lens[rank * per:(rank + 1) * per].tolist()
    cpart = np.frombuffer(b"".join(c for c, _ in part), dtype=np.uint8)
    ppart = np.zeros(int(align16(plain_lens[rank * per:(rank + 1) * per]).sum()), dtype=np.uint8)
    o = 0
    for _, pl in part:
        ppart[o:o + len(pl)] = np.frombuffer(pl, dtype=np.uint8)
        o += int(align16(len(pl)))
    if world > 1:
        lens_all = [torch.empty(per, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(lens_all, torch.from_numpy(comp_lens_part).to(dev))
        comp_lens = torch.cat(lens_all).cpu().numpy()

        def gather_blob(mine, sizes):
            mx = int(max(sizes))
            buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
            buf[:len(mine)] = torch.from_numpy(mine).to(dev)
            outl = [torch.empty(mx, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(outl, buf)
            return torch.cat([outl[r][:int(sizes[r])] for r in range(world)])
        d_comp = gather_blob(cpart, [int(comp_lens[r * per:(r + 1) * per].sum()) for r in range(world)])
        d_plain = gather_blob(ppart, [int(align16(plain_lens[r * per:(r + 1) * per]).sum()) for r in range(world)])
    else:
        comp_lens = comp_lens_part
        d_comp = torch.from_numpy(cpart).to(dev)
        d_plain = torch.from_numpy(ppart).to(dev)
    comp_period, plain_period = int(comp_lens.sum()), int(align16(plain_lens).sum())
    assert d_comp.numel() == comp_period and d_plain.numel() == plain_period

    in_off_g, out_off_g, ranges = ns_layout(comp_lens, plain_lens, total_streams, world)
    lo, hi = ranges[rank]
    n = hi - lo
    in_lo, in_hi, out_lo, out_hi = int(in_off_g[lo]), int(in_off_g[hi]), int(out_off_g[lo]), int(out_off_g[hi])
    in_bytes, out_bytes = in_hi - in_lo, int(np.tile(plain_lens, -(-total_streams // distinct))[lo:hi].sum())
    total_in, total_out = int(in_off_g[-1]), int(np.tile(plain_lens, -(-total_streams // distinct))[:total_streams].sum())
    in_off = np.ascontiguousarray(in_off_g[lo:hi + 1] - np.uint64(in_lo))
    out_off = np.ascontiguousarray(out_off_g[lo:hi + 1] - np.uint64(out_lo))

    # ---- resident shards.  With several ranks rank 0 also holds the WHOLE batch (the `sharded` runs start and end there);
    # its own shard is then a slice of those blobs.
    do_sharded = world > 1 and not a.no_sharded
    full_in = full_out = None
    if do_sharded and rank == 0:
        full_in = tile_slice(torch, d_comp, 0, total_in, pad=256)
        full_out = torch.zeros(int(out_off_g[-1]) + 256, dtype=torch.uint8, device=dev)
        assert in_lo % 16 == 0 and out_lo % 16 == 0
        d_in, d_out = full_in[in_lo:], full_out[out_lo:out_hi + 16]
    else:
        d_in = tile_slice(torch, d_comp, in_lo, in_hi, pad=256)
        d_out = torch.zeros(out_hi - out_lo + 16, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream()
    sptr = C.c_void_p(stream.cuda_stream)
    opt = _native.make_options()
    batch = C.c_void_p()
    rc = lib.lzb_batch_prepare(ctx.handle, _native.FMT_LZMA2, C.byref(opt), d_in.data_ptr(), in_off.ctypes.data, n,
                               d_out.data_ptr(), out_off.ctypes.data, C.byref(batch))
    assert rc == 0, (rc, ctx.last_error())
    kernels_per_step = lib.lzb_batch_kernels_per_launch(batch)
    kernel_name = lib.lzb_batch_kernel_name(batch).decode()  # the K1 variant the planner picked for this batch
    out_len = np.zeros(n, dtype=np.uint64)
    consumed = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=_native.STATUS_DTYPE)
    want_len = np.tile(plain_lens, -(-total_streams // distinct))[lo:hi].astype(np.uint64)

    def check_device(what):
        r = lib.lzb_batch_collect(batch, sptr, out_len.ctypes.data, consumed.ctypes.data, st.ctypes.data)
        assert r == 0 and (st["code"] == 0).all() and (out_len == want_len).all(), f"{what}: decode failed on some stream"
        assert tile_equal(torch, d_out, d_plain, out_lo, out_hi), f"{what}: output differs from the plaintexts"

    def step():
        r = lib.lzb_batch_launch(batch, sptr)
        assert r == 0, (r, ctx.last_error())

    # ---- (i) kernel-only
    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms = []
    for i in range(a.steps):
        with torch.cuda.stream(stream):
            d_out.zero_()  # this step must write every byte it is credited for
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        stream.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        if not a.no_verify:
            check_device(f"timed step {i}")
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = allmax(float(np.sum(step_ms)))
    ms_per_step = total_ms / a.steps
    value = total_out / (ms_per_step * 1e-3) / 1e9
    verified = "skipped" if a.no_verify else (f"every timed step: all {n} streams of the rank's shard byte-identical to their "
                                              "plaintexts (device compare, output zeroed before the step)")
    lib.lzb_batch_destroy(batch)

    # ---- (ii) e2e: pinned host buffers through lzb_decode_batch, every rank its own shard over its own PCIe link
    h_comp = d_comp.cpu()
    h_plain = d_plain.cpu()
    h_in = tile_slice(torch, h_comp, in_lo, in_hi, pad=-256)
    h_out = torch.zeros(out_hi - out_lo + 16, dtype=torch.uint8, pin_memory=True)
    e_len = np.zeros(n, dtype=np.uint64)
    e_cons = np.zeros(n, dtype=np.uint64)
    e_st = np.zeros(n, dtype=_native.STATUS_DTYPE)

    def e2e_step():
        r = lib.lzb_decode_batch(ctx.handle, _native.FMT_LZMA2, C.byref(opt), h_in.data_ptr(), in_off.ctypes.data, n,
                                 h_out.data_ptr(), out_off.ctypes.data, e_len.ctypes.data, e_cons.ctypes.data,
                                 e_st.ctypes.data)
        assert r == 0, (r, ctx.last_error())

    e_steps = max(2, min(a.steps, 3))
    e2e_step()
    h_out.zero_()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = allmax((time.perf_counter() - t0) / e_steps * 1e3)
    assert (e_st["code"] == 0).all() and (e_len == want_len).all()
    if not a.no_verify:
        assert tile_equal(torch, h_out, h_plain, out_lo, out_hi), "e2e: host output differs from the plaintexts"
    e2e_value = total_out / (e2e_ms * 1e-3) / 1e9
    del h_in, h_out

    # ---- (iii) the whole batch starts and ends in rank 0's HBM
    sharded = None
    if do_sharded:
        sharded = ns_sharded(a, torch, dist, lib, ctx, sharding, rank, world, dev, full_in, full_out, in_off_g, out_off_g,
                             ranges, d_plain, total_out, plain_lens, distinct, barrier, allmax)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kernel_ms = float(np.mean(step_ms))
        achieved = (in_bytes + out_bytes) / (kernel_ms * 1e-3) / 1e9
        if world == 1:
            k = min(a.cpu_sample, distinct)
            cpu_obj, cpu_lz = cpu_baselines([c for c, _ in part[:k]], [p for _, p in part[:k]], k, total_streams,
                                            CONFIGS["ns"]["dict_size"])
        else:
            cpu_obj = {"value": None, "unit": "GB/s", "cores": 0, "kind": "port",
                       "sample": "not timed at N > 1 (rank 0 at N = 1 only): see the N = 1 line of the same box"}
            cpu_lz = None
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r02_k1_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("config") == "ns" and tj.get("streams") == n:
                traffic, traffic_src = tj["dram_bytes_per_launch"], "committed profile " + os.path.basename(tpath)
        line = {
            "metric": "decompressed GB/s (batch of independent LZMA2 streams)", "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/u16/u32 integer",
            "data": "synthetic",
            "config": ns_config(total_streams, distinct),
            "detail": {"compressed_bytes_total": total_in, "decompressed_bytes_total": total_out,
                       "streams_rank0": n, "compressed_bytes_rank0": in_bytes, "decompressed_bytes_rank0": out_bytes,
                       "l2": f"inputs+outputs {(in_bytes + out_bytes) >> 20} MiB per GPU and step > {L2_BYTES // 1000000} MB L2 (no flush needed)",
                       "parallelism": f"{world} ranks x contiguous stream range with 1/{world} of the compressed bytes, no "
                                      "data-path collective", "verified": verified,
                       "corpus_generation_s": round(t_gen, 1)},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": total_in, "d2h_bytes_per_step": total_out,
                    "ms_per_step": e2e_ms, "steps": e_steps,
                    "api": "lzb_decode_batch (C ABI) with pinned host buffers, one call per rank on its shard: per-device "
                           "H2D behind K1's input gate, " +
                           ("finished streams moved to the host buffer by the copy engine while K1 decodes (drain mode)"
                            if total_streams // world > 148 * 28 else "output pages stored to the host buffer by K1"),
                    "verified": "every stream of the last step's host output byte-identical (buffer zeroed before)"},
            "gpu_launches": kernels_per_step * a.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": kernel_name,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": in_bytes + out_bytes,
                         "kernel_ms": kernel_ms},
            "cpu_baseline": cpu_obj,
            "clocks": clocks,
        }
        if cpu_lz is not None:
            line["cpu_baseline_liblzma"] = cpu_lz
        if sharded is not None:
            line["sharded"] = sharded
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ns_sharded(a, torch, dist, lib, ctx, sharding, rank, world, dev, full_in, full_out, in_off_g, out_off_g, ranges,
               d_plain, total_out, plain_lens, distinct, barrier, allmax):
    """The whole batch starts and ends in rank 0's HBM: (a) CUDA-IPC peer form, one launch per rank; (b) plain NCCL
    send/recv scatter + decode + gather.  Returns the `sharded` object (rank 0) or None."""
    from lzma_rs_b200 import _native
    lo, hi = ranges[rank]
    n = hi - lo
    total_streams = len(in_off_g) - 1
    opt = _native.make_options()
    res = {}
    in_off = np.ascontiguousarray(in_off_g[lo:hi + 1])
    out_off = np.ascontiguousarray(out_off_g[lo:hi + 1])
    out_len = np.zeros(n, dtype=np.uint64)
    cons = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=_native.STATUS_DTYPE)
    want_len = np.tile(plain_lens, -(-total_streams // distinct))[lo:hi].astype(np.uint64)

    # ---- (a) peers map rank 0's blobs and run scatter + decode + gather as one launch
    handles = [None, None]
    if rank == 0:
        hs = []
        for t in (full_in, full_out):
            h = _native.IpcHandle()
            rc = lib.lzb_ipc_export(ctx.handle, t.data_ptr(), t.numel(), C.byref(h))
            assert rc == 0, (rc, ctx.last_error())
            hs.append(bytes(h))
        handles = hs
    dist.broadcast_object_list(handles, src=0)
    if rank == 0:
        p_in, p_out = full_in.data_ptr(), full_out.data_ptr()
    else:
        ptrs = []
        for raw in handles:
            h = _native.IpcHandle.from_buffer_copy(raw)
            p = C.c_void_p()
            rc = lib.lzb_ipc_open(ctx.handle, C.byref(h), C.byref(p))
            assert rc == 0, (rc, ctx.last_error())
            ptrs.append(p.value)
        p_in, p_out = ptrs

    def p2p_step():
        if rank == 0:  # the owner decodes its own range in place
            r = lib.lzb_decode_batch_device(ctx.handle, _native.FMT_LZMA2, C.byref(opt), p_in, in_off.ctypes.data, n, p_out,
                                            out_off.ctypes.data, out_len.ctypes.data, cons.ctypes.data, st.ctypes.data, None)
        else:
            r = lib.lzb_decode_batch_peer(ctx.handle, _native.FMT_LZMA2, C.byref(opt), p_in, in_off.ctypes.data, n, p_out,
                                          out_off.ctypes.data, out_len.ctypes.data, cons.ctypes.data, st.ctypes.data)
        assert r == 0, (r, ctx.last_error())
        assert (st["code"] == 0).all() and (out_len == want_len).all()

    times = []
    s_steps = max(2, min(a.steps, 3))
    for it in range(1 + s_steps):
        if rank == 0:
            full_out.zero_()
        barrier()
        t0 = time.perf_counter()
        p2p_step()
        barrier()  # the job ends when every rank's pages have landed in rank 0's blob
        dt = allmax(time.perf_counter() - t0)
        if it:
            times.append(dt)
        if rank == 0 and not a.no_verify:
            assert tile_equal(torch, full_out, d_plain, 0, int(out_off_g[-1])), "sharded p2p: rank 0's output blob differs"
    if rank != 0:
        lib.lzb_ipc_close(ctx.handle, p_in)
        lib.lzb_ipc_close(ctx.handle, p_out)
    ms = float(np.mean(times)) * 1e3
    res["p2p_fused"] = {"value": total_out / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms, "steps": s_steps,
                        "api": "lzb_ipc_export/open + lzb_decode_batch_peer (ranks > 0), lzb_decode_batch_device (rank 0)",
                        "verified": "rank 0's whole output blob byte-identical after every step (zeroed before)",
                        "timing": "host clock around the blocking calls, cuda-synchronised, barrier on both sides, max over ranks"}

    # ---- (b) plain NCCL scatter / gather (no overlap): the comparison point
    fn = sharding.cuda_decode_fn(ctx, 1)
    caps = np.tile(plain_lens, -(-total_streams // distinct))[:total_streams] if rank == 0 else None
    times = []
    out = None
    for it in range(2):
        out = None
        barrier()
        t0 = time.perf_counter()
        out = sharding.decode_sharded_tensors(fn, full_in[:int(in_off_g[-1]) + 16] if rank == 0 else None,
                                              in_off_g if rank == 0 else None, caps, src=0)
        barrier()
        dt = allmax(time.perf_counter() - t0)
        if it:
            times.append(dt)
    if rank == 0:
        out_t, _, lens, codes = out
        assert (codes == 0).all()
        if not a.no_verify:
            assert tile_equal(torch, out_t, d_plain, 0, int(out_off_g[-1])), "sharded nccl: output differs"
    ms = float(np.mean(times)) * 1e3
    res["nccl_scatter_gather"] = {"value": total_out / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms, "steps": 1,
                                  "api": "lzma_rs_b200.sharding.decode_sharded_tensors (NCCL send/recv, unoverlapped)"}
    return res if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="ns", choices=sorted(CONFIGS), help="ns = the benchmark line; others informational")
    ap.add_argument("--streams", type=int, default=0, help="streams (ns: in the whole batch; others: per GPU; 0 = the config's own)")
    ap.add_argument("--no-sharded", action="store_true", help="ns, N > 1: skip the rank-0-resident scatter/decode/gather runs")
    ap.add_argument("--distinct", type=int, default=0, help="distinct streams to generate (0 = auto); the rest are tiled")
    ap.add_argument("--cpu-sample", type=int, default=0, help="streams of the workload the CPU baseline decodes (0 = auto)")
    ap.add_argument("--no-verify", action="store_true")
    a = ap.parse_args()
    assert a.warmup >= 3 or a.impl == "reference", "timing rules: at least 3 warm-up steps"
    global CFG
    CFG = CONFIGS[a.config]
    a.streams = a.streams or CFG["streams"]
    a.cpu_sample = a.cpu_sample or (1024 if a.config == "ns" else 4096)  # ns: 1 024 streams = 0.36 GB out per pass

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncpu = os.cpu_count() or 1
    workers = max(1, min(32, ncpu // max(1, world)))
    distinct = a.distinct or (min(a.streams, 4096 if CFG["kind"] == "mixed" and CFG["stream_bytes"] <= 65536 else 1024)
                              if ncpu >= 16 else min(a.streams, 1024))
    if CFG["kind"] in ("rep0", "stored"):
        distinct = min(distinct, 256)

    workload = CFG["name"].format(n=a.streams, d=min(a.distinct or NS_DISTINCT, a.streams))

    if a.config == "ns" and a.impl == "ours":
        return bench_ns(a, rank, local_rank, world)

    # ---------------------------------------------------------------- reference arm: CPU path only
    if a.impl == "reference":
        if rank != 0:
            return
        if a.config == "ns":  # a bounded sample of the SAME batch: its first distinct streams
            k = min(a.cpu_sample, a.streams, a.distinct or NS_DISTINCT)
            part = ns_generate(0, k, workers)
            comp, plain = [c for c, _ in part], [p for _, p in part]
        else:
            comp, plain = build_corpus(0, min(a.streams, max(a.cpu_sample, 64)), distinct, workers)
        ncap, quota = effective_cpus()
        gbs, dt, k, total, used = cpu_reference_run(comp, plain, a.steps, a.warmup, cpu_thread_candidates(ncap, quota),
                                                    a.cpu_sample)
        line = {"impl": "reference", "metric": "decompressed GB/s (batch of independent LZMA2 streams)", "value": gbs,
                "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "strong" if a.config == "ns" else "weak", "vs_baseline": None,
                "dtype": "u8/u16/u32 integer",
                "data": "synthetic",
                "config": ns_config(a.streams, min(a.distinct or NS_DISTINCT, a.streams)) if a.config == "ns" else
                          {"workload": workload, "sample": f"{k} streams of the workload per step"},
                "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": effective_cores(used, ncap, quota), "kind": "port",
                                 "per_core_gbs": gbs / effective_cores(used, ncap, quota),
                                 "sample": f"{k} of {a.streams} streams ({total} B out) per step, one stream per task, "
                                           f"best of thread counts {cpu_thread_candidates(ncap, quota)} -> {used} pthreads "
                                           f"(os.cpu_count {ncpu}, cgroup quota {quota}); C restatement of lzma-rs "
                                           "src/decode (reference is Rust, no rustc in the image)"},
                "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm
    comp, plain = build_corpus(rank, a.streams, distinct, workers)  # before CUDA init (fork pool)

    import torch
    import torch.distributed as dist
    from lzma_rs_b200 import Context, _native
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _native.load()
    ctx = Context(local_rank)

    if CFG["kind"] == "xz":
        return bench_xz(a, lib, ctx, comp, plain, workload, world, rank)

    n = len(comp)
    blob, in_off = _native.pack_streams(comp)
    sizes = np.array([len(p) for p in plain], dtype=np.uint64)
    caps = (sizes + np.uint64(15)) // np.uint64(16) * np.uint64(16)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(caps, out=out_off[1:])
    in_bytes, out_bytes = int(in_off[-1]), int(sizes.sum())

    d_in = torch.from_numpy(blob).cuda()
    d_out = torch.zeros(int(out_off[-1]) + 16, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()  # an explicit stream: the library treats a NULL handle as "use the ctx stream"
    torch.cuda.synchronize()
    sptr = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0
    opt = _native.make_options()
    batch = C.c_void_p()
    rc = lib.lzb_batch_prepare(ctx.handle, _native.FMT_LZMA2, C.byref(opt), d_in.data_ptr(), in_off.ctypes.data, n,
                               d_out.data_ptr(), out_off.ctypes.data, C.byref(batch))
    assert rc == 0, (rc, ctx.last_error())
    kernels_per_step = lib.lzb_batch_kernels_per_launch(batch)
    kernel_name = lib.lzb_batch_kernel_name(batch).decode()  # the K1 variant the planner picked for this batch

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        r = lib.lzb_batch_launch(batch, sptr)
        assert r == 0, (r, ctx.last_error())

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record(stream)
    for i in range(a.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])

    out_len = np.zeros(n, dtype=np.uint64)
    consumed = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=_native.STATUS_DTYPE)
    rc = lib.lzb_batch_collect(batch, sptr, out_len.ctypes.data, consumed.ctypes.data, st.ctypes.data)
    assert rc == 0 and (st["code"] == 0).all(), "decode failed on some stream"
    verified = "skipped"
    if not a.no_verify:  # bit-exactness of EVERY stream of the timed batch against the plaintexts it was made from
        host = d_out.cpu().numpy()
        for i in range(n):
            o = int(out_off[i])
            assert host[o:o + int(out_len[i])].tobytes() == plain[i], f"stream {i} differs"
        verified = f"all {n} streams byte-identical to their plaintexts"

    # ---------------- e2e: host buffers through the reference-facing C-ABI call
    h_in = torch.from_numpy(blob).pin_memory()
    h_out = torch.empty(int(out_off[-1]) + 16, dtype=torch.uint8).pin_memory()
    e_len = np.zeros(n, dtype=np.uint64)
    e_cons = np.zeros(n, dtype=np.uint64)
    e_st = np.zeros(n, dtype=_native.STATUS_DTYPE)

    def e2e_step():
        r = lib.lzb_decode_batch(ctx.handle, _native.FMT_LZMA2, C.byref(opt), h_in.data_ptr(), in_off.ctypes.data, n,
                                 h_out.data_ptr(), out_off.ctypes.data, e_len.ctypes.data, e_cons.ctypes.data,
                                 e_st.ctypes.data)
        assert r == 0, (r, ctx.last_error())

    e_steps = max(3, min(a.steps, 5))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e_steps
    assert (e_st["code"] == 0).all()
    if not a.no_verify:
        hv = h_out.numpy()
        for i in range(0, n, 97):
            o = int(out_off[i])
            assert hv[o:o + int(e_len[i])].tobytes() == plain[i]

    # ---------------- max over ranks
    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    ms_per_step = total_ms_max / a.steps
    value = world * out_bytes / (ms_per_step * 1e-3) / 1e9
    e2e_value = world * out_bytes / (e2e_ms_max * 1e-3) / 1e9

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kernel_ms = float(np.mean(step_ms))  # this rank's average launch duration, CUDA events on the launch stream
        achieved = (in_bytes + out_bytes) / (kernel_ms * 1e-3) / 1e9
        ncap, quota = effective_cpus()
        if world == 1:
            cpu_gbs, cpu_dt, cpu_k, cpu_total, cpu_used = cpu_reference_run(comp, plain, 2, 1,
                                                                            cpu_thread_candidates(ncap, quota), a.cpu_sample)
            cpu_obj = {"value": cpu_gbs, "unit": "GB/s", "cores": effective_cores(cpu_used, ncap, quota), "kind": "port",
                       "sample": f"{cpu_k} of {n} streams ({cpu_total} B out), one stream per task, best of thread "
                                 f"counts {cpu_thread_candidates(ncap, quota)} -> {cpu_used} pthreads (os.cpu_count "
                                 f"{ncpu}, cgroup quota {quota}); C restatement of lzma-rs src/decode (oracle/)"}
        else:  # the CPU baseline is a property of the box, measured at N=1 only: here the other ranks share its cores
            cpu_obj = {"value": None, "unit": "GB/s", "cores": 0, "kind": "port",
                       "sample": "not timed at N > 1 (rank 0 at N = 1 only): see the N = 1 line of the same box"}
        traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch, from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "r01_k1_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("streams") == n and a.config == "c2":
                traffic = tj["dram_bytes_per_launch"]
        line = {
            "metric": "decompressed GB/s (batch of independent LZMA2 streams)", "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u16/u32 integer",
            "data": "synthetic",
            "config": {"workload": workload, "streams_per_gpu": n, "distinct_streams": min(distinct, n),
                       "compressed_bytes_per_gpu": in_bytes, "decompressed_bytes_per_gpu": out_bytes,
                       "l2": (f"inputs+outputs {(in_bytes + out_bytes) >> 20} MiB per step > {L2_BYTES // 1000000} MB L2 "
                              "(no flush needed)") if in_bytes + out_bytes > L2_BYTES else
                             (f"inputs+outputs {(in_bytes + out_bytes) >> 20} MiB per step FIT the {L2_BYTES // 1000000} MB L2 and "
                              "are not flushed: reduced --streams run, not a benchmark configuration"),
                       "parallelism": f"{world} x independent shard (no collective)", "verified": verified},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": in_bytes * world,
                    "d2h_bytes_per_step": out_bytes * world, "ms_per_step": e2e_ms_max,
                    "api": "lzb_decode_batch (C ABI) with pinned host buffers"},
            "gpu_launches": kernels_per_step * a.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": kernel_name,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": in_bytes + out_bytes, "kernel_ms": kernel_ms},
            "cpu_baseline": cpu_obj,
            "clocks": clocks,
        }
        print(json.dumps(line))
    lib.lzb_batch_destroy(batch)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
