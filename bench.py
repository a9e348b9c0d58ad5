#!/usr/bin/env python3
"""bench.py -- decompressed GB/s of the many-stream LZMA2 decode path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--streams S] [--distinct D]

Workload at N=1: BASELINE.json configs[1] -- 4 096 independent 64 KiB raw-LZMA2 streams (lc3 lp0 pb2, dict 256 KiB,
seeded "mixed literal/match" text, liblzma preset 6).  For N>1 every rank decodes its own
batch of that shape (weak scaling, no data-path collective: streams are independent).

A step = one pass of the hot path (K1, lzb_decode_kernel) over the whole batch.
  value : kernel-only, inputs and outputs resident in HBM, CUDA events on the launching stream, max over ranks.
  e2e   : the same batch through the reference-facing C-ABI call lzb_decode_batch with pinned HOST buffers
          (H2D copy of the compressed streams + scan + decode + D2H copy of the output inside the timed region).
  roofline     : algorithmic bytes (compressed read once + decompressed written once) / kernel time vs the measured
                 HBM copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline : the C oracle (line-by-line restatement of lzma-rs's src/decode, oracle/) on the host cores.
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

# BASELINE.json configs that are raw-LZMA2 batches.  c2 is the benchmark line (configs[1], the
# largest single-GPU configuration the metric is quoted on); c3 / c5 are informational (`--config`).
CONFIGS = {
    "c2": dict(index=2, stream_bytes=65536, dict_size=1 << 18, streams=4096, kind="mixed",
               name="C2: {n} independent raw LZMA2 streams x 65536 B (lc3 lp0 pb2, dict 256 KiB, seeded mixed "
                    "literal/match text, liblzma preset 6) per GPU"),
    "c3": dict(index=3, stream_bytes=262144, dict_size=1 << 20, streams=8192, kind="mixed",
               name="C3 shard: {n} independent raw LZMA2 streams x 262144 B (dict 1 MiB, 3 chunks each) per GPU"),
    "c5": dict(index=5, stream_bytes=262144, dict_size=1 << 20, streams=8192, kind="rep0",
               name="C5: {n} raw LZMA2 streams x 262144 B of all-overlapping rep0 matches (dist=1, len=273) per GPU"),
    "ns": dict(index=6, stream_bytes=0, dict_size=1 << 20, streams=8192, kind="mixed",
               name="NS shard: {n} independent raw LZMA2 streams, sizes log-uniform in [64 KiB, 1 MiB], dict 1 MiB, per GPU "
                    "(north-star sweep: 65 536 streams over 8 GPUs)"),
    "c6": dict(index=7, stream_bytes=1 << 20, dict_size=1 << 20, streams=2048, kind="stored",
               name="C6: {n} raw LZMA2 streams x 1 MiB of stored chunks only (16 x `01 FF FF` + 64 KiB, the only LZMA2 the "
                    "reference's own encoder writes, src/encode/lzma2.rs:4-26) per GPU -- the byte-bound end of the path"),
    "c4": dict(index=4, stream_bytes=1 << 20, dict_size=1 << 20, streams=1024, kind="xz",
               name="C4: {n} .xz files x 1 MiB (4 blocks of 256 KiB each, LZMA2 filter, CRC32 block check) per GPU, "
                    "host API only (container walk on the host, K1 decode + K3 CRC on the GPU)"),
}
CFG = CONFIGS["c2"]


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="c2 = the benchmark line; others informational")
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (0 = the config's own count)")
    ap.add_argument("--distinct", type=int, default=0, help="distinct streams to generate (0 = auto); the rest are tiled")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="streams of the workload the CPU baseline decodes")
    ap.add_argument("--no-verify", action="store_true")
    a = ap.parse_args()
    assert a.warmup >= 3 or a.impl == "reference", "timing rules: at least 3 warm-up steps"
    global CFG
    CFG = CONFIGS[a.config]
    a.streams = a.streams or CFG["streams"]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncpu = os.cpu_count() or 1
    workers = max(1, min(32, ncpu // max(1, world)))
    distinct = a.distinct or (min(a.streams, 4096 if CFG["kind"] == "mixed" and CFG["stream_bytes"] <= 65536 else 1024)
                              if ncpu >= 16 else min(a.streams, 1024))
    if CFG["kind"] in ("rep0", "stored"):
        distinct = min(distinct, 256)

    workload = CFG["name"].format(n=a.streams)

    # ---------------------------------------------------------------- reference arm: CPU path only
    if a.impl == "reference":
        if rank != 0:
            return
        comp, plain = build_corpus(0, min(a.streams, max(a.cpu_sample, 64)), distinct, workers)
        ncap, quota = effective_cpus()
        gbs, dt, k, total, used = cpu_reference_run(comp, plain, a.steps, a.warmup, cpu_thread_candidates(ncap, quota),
                                                    a.cpu_sample)
        line = {"impl": "reference", "metric": "decompressed GB/s (batch of independent LZMA2 streams)", "value": gbs,
                "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u16/u32 integer",
                "data": "synthetic", "config": {"workload": workload, "sample": f"{k} streams of the workload per step"},
                "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": effective_cores(used, ncap, quota), "kind": "port",
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
                         "kernel": {"rep0": "lzb_decode_fill_kernel", "stored": "lzb_stored_decode_kernel"}.get(CFG["kind"], "lzb_decode_kernel"),
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
