#!/usr/bin/env python3
"""tools/kbench.py -- kernel experiments: times K1 of several builds of the library on ONE corpus in ONE process.

    python tools/kbench.py [--config c2] [--steps 5] build/variants/a.so build/variants/b.so ...

Every build decodes the same device-resident batch through lzb_batch_prepare / lzb_batch_launch (CUDA events on the
launch stream, 3 warm-ups); its output is compared byte-for-byte with the plaintexts.  Prints one line per build.
Development tool only: the numbers quoted anywhere come from bench.py.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    bench.CFG = bench.CONFIGS[a.config]
    n = a.streams or bench.CFG["streams"]
    distinct = min(n, 256 if bench.CFG["kind"] in ("rep0", "stored") else (4096 if bench.CFG["stream_bytes"] <= 65536 else 1024))
    comp, plain = bench.build_corpus(0, n, distinct, max(1, min(32, os.cpu_count() or 1)))
    import torch
    from lzma_rs_b200 import _native
    blob, in_off = _native.pack_streams(comp)
    sizes = np.array([len(p) for p in plain], dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((sizes + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
    d_in = torch.from_numpy(blob).cuda()
    want = torch.zeros(int(out_off[-1]) + 16, dtype=torch.uint8)
    wv = want.numpy()
    for i in range(n):
        wv[int(out_off[i]):int(out_off[i]) + len(plain[i])] = np.frombuffer(plain[i], dtype=np.uint8)
    d_want = want.cuda()
    out_bytes = int(sizes.sum())
    stream = torch.cuda.Stream()
    sptr = C.c_void_p(stream.cuda_stream)
    opt = _native.make_options()
    for spec in a.libs:  # path[@NAME=value[,NAME=value]]: environment switches of the library for this build only
        path, _, envs = spec.partition("@")
        env = dict(kv.split("=", 1) for kv in envs.split(",") if kv)
        os.environ.update(env)
        lib = _native.bind(os.path.abspath(path))
        h = C.c_void_p()
        assert lib.lzb_create(C.byref(h), 0) == 0
        d_out = torch.zeros(int(out_off[-1]) + 16, dtype=torch.uint8, device="cuda")
        batch = C.c_void_p()
        rc = lib.lzb_batch_prepare(h, _native.FMT_LZMA2, C.byref(opt), d_in.data_ptr(), in_off.ctypes.data, n,
                                   d_out.data_ptr(), out_off.ctypes.data, C.byref(batch))
        assert rc == 0, rc
        for _ in range(3):
            assert lib.lzb_batch_launch(batch, sptr) == 0
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        ev[0].record(stream)
        for i in range(a.steps):
            assert lib.lzb_batch_launch(batch, sptr) == 0
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps))
        st = np.zeros(n, dtype=_native.STATUS_DTYPE)
        ol = np.zeros(n, dtype=np.uint64)
        cs = np.zeros(n, dtype=np.uint64)
        assert lib.lzb_batch_collect(batch, sptr, ol.ctypes.data, cs.ctypes.data, st.ctypes.data) == 0
        ok = bool((st["code"] == 0).all()) and bool((ol == sizes).all()) and bool(torch.equal(d_out, d_want))
        med = ms[len(ms) // 2]
        for k in env:
            del os.environ[k]
        print(f"{os.path.basename(spec):28s} median {med:9.3f} ms  min {ms[0]:9.3f}  {out_bytes / med / 1e6:9.2f} GB/s  "
              f"bit-exact={'yes' if ok else 'NO'}", flush=True)
        lib.lzb_batch_destroy(batch)
        lib.lzb_destroy(h)
        del d_out


if __name__ == "__main__":
    main()
