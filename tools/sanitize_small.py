#!/usr/bin/env python3
"""tools/sanitize_small.py -- a small workload that touches every kernel added late in round 1, sized for
compute-sanitizer (memcheck / racecheck):  K4 store + frame, K5 literal encoder, K6 stored-chunk decode, K1 behind
the input gate (pinned output, > 16 MiB of input), K1 with a placement plan.  Every result is checked against the oracle.

    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import corpus
    import gpu_util
    import oracle_py as oracle
    from lzma_rs_b200 import Context
    ctx = Context()
    rng = np.random.default_rng(1)
    datas = [b"", b"a", rng.bytes(65536), rng.bytes(65537), corpus.mixed_text(5, 70_000), rng.bytes(200_001)]
    assert ctx.encode_batch(1, datas) == [oracle.lzma2_compress(d) for d in datas]
    assert ctx.encode_batch(2, datas) == [oracle.xz_compress(d) for d in datas]
    small = [b"", b"a", corpus.mixed_text(6, 3000), rng.bytes(2000)] * 5
    assert ctx.encode_batch(0, small) == [oracle.lzma_compress(d) for d in small]
    # K6 (device path routes stored-only streams to the copy kernel) next to K1 in one batch
    streams = [corpus.stored_lzma2(d) for d in datas[1:]] + [corpus.raw_lzma2(corpus.mixed_text(7, 30_000))]
    caps = [len(d) for d in datas[1:]] + [30_000]
    b = gpu_util.DeviceBatch(ctx, 1, streams, caps).decode()
    for i, s in enumerate(streams):
        assert b.st[i]["code"] == 0 and b.output(i) == oracle.lzma2_decompress(s).out, i
    # input gate: pinned output, 20 MiB of (mostly stored) input, a few compressed streams among them
    blocks = [rng.bytes(1 << 20) for _ in range(4)]
    streams = [corpus.stored_lzma2(blocks[i % 4]) for i in range(20)] + [corpus.raw_lzma2(corpus.mixed_text(8 + i, 20_000)) for i in range(6)]
    plains = [blocks[i % 4] for i in range(20)] + [corpus.mixed_text(8 + i, 20_000) for i in range(6)]
    outs, out_len, consumed, st = gpu_util.host_decode_pinned(ctx, 1, streams, [len(p) for p in plains])
    assert (st["code"] == 0).all() and outs == plains
    # placement plan: more streams than resident warps, sizes spread over 64x (tiny streams keep the run short)
    base_plain = [corpus.mixed_text(100 + i, int(256 * 64 ** rng.random())) for i in range(200)]
    base = [corpus.raw_lzma2(p, dict_size=1 << 16) for p in base_plain]
    n = 148 * 28 + 600
    b = gpu_util.DeviceBatch(ctx, 1, [base[i % 200] for i in range(n)], [len(base_plain[i % 200]) for i in range(n)]).decode()
    assert (b.st["code"] == 0).all()
    for i in range(0, n, 37):
        assert b.output(i) == base_plain[i % 200], i
    print("sanitize_small: all checks passed")


if __name__ == "__main__":
    main()
