#!/usr/bin/env python3
"""tools/fuzz_soak.py -- differential fuzz soak in the style of the reference's fuzz targets (fuzz/fuzz_targets/
compare_xz.rs:28-37: both fail the same way or both decode the same bytes), against the oracle.

    python tools/fuzz_soak.py --backend emul --rounds 50 --seed 1      K1's source compiled for the CPU (no GPU needed)
    python tools/fuzz_soak.py --backend gpu  --rounds 50 --seed 1      the CUDA path through the C ABI

Each round mutates a fresh set of seed streams (all three formats, several lc/lp/pb, stored chunks, multi-chunk
LZMA2, multi-block / chained .xz) 400 times per format and compares display string, output bytes and consumed count."""
import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def seeds_for(rnd, corpus):
    r = lambda a, b: rnd.randrange(a, b)  # noqa: E731
    props = [(3, 0, 2), (0, 0, 0), (4, 0, 4), (2, 2, 1), (0, 4, 0), (1, 3, 3)]
    out = {0: [], 1: [], 2: []}
    for _ in range(4):
        lc, lp, pb = rnd.choice(props)
        out[0].append(corpus.lzma_alone(corpus.mixed_text(r(0, 1 << 30), r(1, 30_000)), dict_size=rnd.choice([4096, 1 << 16]),
                                        lc=lc, lp=lp, pb=pb))
        out[1].append(corpus.raw_lzma2(corpus.mixed_text(r(0, 1 << 30), r(1, 300_000)), dict_size=rnd.choice([4096, 1 << 20]),
                                       lc=lc, lp=lp, pb=pb))
    out[0].append(corpus.lzma_alone_known_size(corpus.mixed_text(r(0, 1 << 30), r(1, 9000)), dict_size=4096))
    out[0].append(corpus.dumb_lzma(corpus.mixed_text(r(0, 1 << 30), r(0, 600))))
    out[1].append(corpus.stored_lzma2(corpus.mixed_text(r(0, 1 << 30), r(0, 140_000))))
    out[1].append(corpus.raw_lzma2(bytes(r(1, 100_000)), dict_size=1 << 16))
    out[2].append(corpus.xz_file(corpus.mixed_text(r(0, 1 << 30), r(1, 20_000)), block_size=r(500, 6000),
                                 check=rnd.choice([corpus.CHECK_CRC32, corpus.CHECK_CRC64])))
    out[2].append(corpus.xz_file(corpus.mixed_text(r(0, 1 << 30), r(1, 60_000)), block_size=1 << 14, check=corpus.CHECK_CRC64,
                                 with_sizes=True))
    out[2].append(corpus.xz_file(b"", check=corpus.CHECK_CRC32))
    return out


def mutate(rnd, b):
    b = bytearray(b)
    if not b:
        return bytes([rnd.randrange(256)])
    k = rnd.random()
    if k < 0.5:
        for _ in range(rnd.choice([1, 1, 1, 2, 4, 9])):
            b[rnd.randrange(len(b))] = rnd.randrange(256)
    elif k < 0.65:
        b = b[:rnd.randrange(len(b))]
    elif k < 0.75:
        i = rnd.randrange(len(b))
        b[i:i] = bytes(rnd.randrange(256) for _ in range(rnd.choice([1, 2, 7, 40])))
    elif k < 0.85:
        i = rnd.randrange(len(b))
        del b[i:i + rnd.choice([1, 2, 5, 33])]
    elif k < 0.95:
        i = rnd.randrange(min(len(b), 32))
        b[i] ^= 1 << rnd.randrange(8)
    else:  # splice the tail of the stream onto an earlier point
        i, j = sorted((rnd.randrange(len(b)), rnd.randrange(len(b))))
        b = b[:i] + b[j:]
    return bytes(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["emul", "gpu"], default="emul")
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--per-format", type=int, default=400)
    ap.add_argument("--cases", action="store_true",
                    help="mutate the streams of the parity corpus (tests/cases.py: hand-encoded, chained .xz, ...) instead")
    a = ap.parse_args()
    import corpus
    import parity
    if a.backend == "gpu":
        import gpu_util
        from lzma_rs_b200 import Context
        ctx = Context()
        decode = lambda fmt, streams, opts: gpu_util.host_decode(ctx, fmt, streams, opts)  # noqa: E731
    else:
        import test_emul_parity
        decode = test_emul_parity._decode
    rnd = random.Random(a.seed)
    total, bad_total, t0 = 0, 0, time.time()
    case_seeds = None
    if a.cases:
        import cases
        case_seeds = {0: [], 1: [], 2: []}
        for fam in ("valid_lzma2_cases", "valid_lzma_cases", "hand_encoded_cases", "xz_cases", "xz_chain_cases"):
            for name, fmt, stream, opts in getattr(cases, fam)():
                if not opts and len(stream) < 70_000:
                    case_seeds[fmt].append(stream)
    for rd in range(a.rounds):
        seeds = seeds_for(rnd, corpus) if case_seeds is None else {f: rnd.sample(v, min(len(v), 40)) for f, v in case_seeds.items()}
        for fmt, srcs in seeds.items():
            named = [(f"r{rd}-f{fmt}-{i}", mutate(rnd, srcs[i % len(srcs)])) for i in range(a.per_format)]
            bad = parity.check_group(decode, fmt, {}, named)
            total += len(named)
            if bad:
                bad_total += len(bad)
                print(f"round {rd} fmt {fmt}: {len(bad)} mismatches", flush=True)
                for line in bad[:5]:
                    print("   " + line, flush=True)
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                names = {b.split(":")[0] for b in bad}
                for name, s in named:
                    if name in names:
                        open(os.path.join(ROOT, "gpurun_out", f"fuzzfail-{a.backend}-s{a.seed}-{name}.bin"), "wb").write(s)
    print(f"fuzz_soak backend={a.backend} seed={a.seed}: {total} mutated streams, {bad_total} mismatches, {time.time() - t0:.0f} s",
          flush=True)
    return 1 if bad_total else 0


if __name__ == "__main__":
    sys.exit(main())
