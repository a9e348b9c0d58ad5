#!/usr/bin/env python3
"""tools/e2e_trace.py -- host-API timeline: runs lzb_decode_batch on the bench corpus with pinned host buffers and
LZB_TRACE=1 (the library prints where the end-to-end time goes: planning, launch, upload, kernel in situ).

    python tools/e2e_trace.py [--config c2] [--steps 4]           (LZB_NO_GATE=1 for the ungated upload)
"""
import argparse
import ctypes as C
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
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    os.environ["LZB_TRACE"] = "1"
    bench.CFG = bench.CONFIGS[a.config]
    n = bench.CFG["streams"]
    comp, plain = bench.build_corpus(0, n, min(n, 4096 if bench.CFG["stream_bytes"] <= 65536 else 1024), min(32, os.cpu_count() or 1))
    import torch
    from lzma_rs_b200 import Context, _native
    lib = _native.load()
    ctx = Context()
    blob, in_off = _native.pack_streams(comp)
    sizes = np.array([len(p) for p in plain], dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((sizes + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
    h_in = torch.from_numpy(blob).pin_memory()
    h_out = torch.empty(int(out_off[-1]) + 16, dtype=torch.uint8).pin_memory()
    ol, cs = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=_native.STATUS_DTYPE)
    opt = _native.make_options()
    fmt = _native.FMT_XZ if bench.CFG["kind"] == "xz" else _native.FMT_LZMA2
    for i in range(a.steps):
        t0 = time.perf_counter()
        rc = lib.lzb_decode_batch(ctx.handle, fmt, C.byref(opt), h_in.data_ptr(), in_off.ctypes.data, n, h_out.data_ptr(),
                                  out_off.ctypes.data, ol.ctypes.data, cs.ctypes.data, st.ctypes.data)
        dt = time.perf_counter() - t0
        assert rc == 0 and (st["code"] == 0).all()
        print(f"step {i}: {dt * 1e3:.3f} ms  {int(sizes.sum()) / dt / 1e9:.2f} GB/s", file=sys.stderr, flush=True)
    hv = h_out.numpy()
    for i in range(0, n, 53):
        o = int(out_off[i])
        assert hv[o:o + int(ol[i])].tobytes() == plain[i]


if __name__ == "__main__":
    main()
