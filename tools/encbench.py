#!/usr/bin/env python3
"""tools/encbench.py -- throughput of the compress-side kernels on device-resident batches (informational).

    python tools/encbench.py            K4 (stored-chunk LZMA2 / XZ): 2048 x 1 MiB;  K5 (literal-only .lzma): 4096 x 64 KiB
Every output is compared with the oracle's encoding of the same plaintext (sampled streams)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run(lib, ctx, fmt, name, datas, steps=5):
    import torch
    import oracle_py as oracle
    from lzma_rs_b200 import _native
    n = len(datas)
    blob, in_off = _native.pack_streams(datas)
    caps = np.array([lib.lzb_encode_bound(fmt, None, len(d)) for d in datas], dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((caps + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
    d_in = torch.from_numpy(blob).cuda()
    d_out = torch.zeros(int(out_off[-1]) + 16, dtype=torch.uint8, device="cuda")
    ol, st = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=_native.STATUS_DTYPE)
    stream = torch.cuda.Stream()

    def step():
        rc = lib.lzb_encode_batch_device(ctx.handle, fmt, None, d_in.data_ptr(), in_off.ctypes.data, n, d_out.data_ptr(),
                                         out_off.ctypes.data, ol.ctypes.data, st.ctypes.data, C.c_void_p(stream.cuda_stream))
        assert rc == 0 and (st["code"] == 0).all()

    for _ in range(3):
        step()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record(stream)
    for i in range(steps):
        step()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))[steps // 2]
    host = d_out.cpu().numpy()
    enc = {0: oracle.lzma_compress, 1: oracle.lzma2_compress, 2: oracle.xz_compress}[fmt]
    for i in range(0, n, max(1, n // 16)):
        o = int(out_off[i])
        assert host[o:o + int(ol[i])].tobytes() == enc(datas[i]), i
    in_bytes, out_bytes = sum(len(d) for d in datas), int(ol.sum())
    print(f"{name:44s} {ms:8.3f} ms per call (incl. plan upload + result fetch)  {in_bytes / ms / 1e6:8.1f} GB/s plaintext  "
          f"{(in_bytes + out_bytes) / ms / 1e6:8.1f} GB/s read+write  bit-exact=yes", flush=True)


def main():
    from lzma_rs_b200 import Context, _native
    lib = _native.load()
    ctx = Context()
    rng = np.random.default_rng(5)
    big = [rng.bytes(1 << 20) for _ in range(64)]
    run(lib, ctx, 1, "K4 lzma2_compress 2048 x 1 MiB", [big[i % 64] for i in range(2048)])
    run(lib, ctx, 2, "K4 xz_compress    2048 x 1 MiB", [big[i % 64] for i in range(2048)])
    import corpus
    small = [corpus.mixed_text(7000 + i, 65536) for i in range(256)]
    run(lib, ctx, 0, "K5 lzma_compress  4096 x 64 KiB", [small[i % 256] for i in range(4096)], steps=3)


if __name__ == "__main__":
    main()
