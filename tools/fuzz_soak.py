#!/usr/bin/env python3
"""tools/fuzz_soak.py -- differential fuzz soak in the style of the reference's fuzz targets (fuzz/fuzz_targets/
compare_xz.rs:28-37: both fail the same way or both decode the same bytes), against the oracle.

    python tools/fuzz_soak.py --backend emul --rounds 50 --seed 1      K1's source compiled for the CPU (no GPU needed)
    python tools/fuzz_soak.py --backend gpu  --rounds 50 --seed 1      the CUDA path through the C ABI (host entry point)
    python tools/fuzz_soak.py --backend gpu-device ...                 the device-resident entry point (not yet run: added
                                                                       after round 1's GPU budget was spent)

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


def xz_structured(rnd, corpus):
    """A .xz file assembled field by field with CORRECT CRCs around deliberately odd values, so that the container walk
    is exercised behind its CRC checks (random byte flips almost always die at the first header CRC): size fields that
    lie, reserved flag bits, unknown / several filters, bad property sizes, non-zero paddings, wrong checks, index
    records that disagree, footers that disagree, trailing bytes."""
    import struct
    import zlib
    r, ch = rnd.random, rnd.choice
    check = ch([0, 1, 1, 4, 4, 0x0A, 2, 0x0F]) if r() < 0.3 else ch([1, 4, 0])
    flags = bytes([0 if r() > 0.02 else rnd.randrange(256), check])
    out = bytearray(b"\xfd7zXZ\0" + flags + struct.pack("<I", zlib.crc32(flags) ^ (1 if r() < 0.02 else 0)))
    records = []
    for _ in range(ch([0, 1, 1, 2, 3])):
        plain = corpus.mixed_text(rnd.randrange(1 << 30), ch([0, 1, 50, 3000, 20_000]))
        pay = corpus.raw_lzma2(plain, dict_size=1 << 16) if r() > 0.2 else corpus.stored_lzma2(plain)
        nf = ch([1, 1, 1, 1, 2, 3, 4])
        for _k in range(nf - 1):
            pay = corpus.raw_lzma2(pay, dict_size=1 << 16)
        with_sizes = r() < 0.4
        bflags = (nf - 1) | (0xC0 if with_sizes else 0)
        if r() < 0.1:
            bflags = (bflags & ~0xC0) | ch([0x40, 0x80])      # only one of the two size fields
        if r() < 0.05:
            bflags |= ch([0x04, 0x08, 0x10, 0x20])            # reserved bits
        body = bytes([bflags])
        if bflags & 0x40:
            body += corpus._multibyte(len(pay) + (0 if r() > 0.15 else ch([-1, 1, 1 << 20])) if len(pay) else 0)
        if bflags & 0x80:
            body += corpus._multibyte(max(0, len(plain) + (0 if r() > 0.15 else ch([-1, 1, 70_000]))))
        for _k in range(nf):
            fid = 0x21 if r() > 0.05 else ch([0x03, 0x00, 0x4000000000000001])
            psz = 1 if r() > 0.05 else ch([0, 2, 5])
            body += corpus._multibyte(fid) + corpus._multibyte(psz) + bytes([0x16] * min(psz, 5))
        total = 1 + len(body) + 4
        padded = (total + 3) & ~3
        if r() < 0.05:
            padded += 4 * ch([1, 2])                          # larger header than needed (legal: zero padding)
        pad = bytes(padded - total) if r() > 0.05 else bytes([ch([0, 1])] * (padded - total))
        hdr = bytes([padded // 4 - 1]) + body + pad
        hdr += struct.pack("<I", zlib.crc32(hdr) ^ (1 if r() < 0.02 else 0))
        blk = hdr + pay
        unpadded = len(blk)
        blk += (b"\0" if r() > 0.05 else b"\1") * ((4 - len(blk) % 4) % 4)
        wrong = r() < 0.05
        if check == 1:
            blk += struct.pack("<I", zlib.crc32(plain) ^ wrong)
            unpadded += 4
        elif check == 4:
            blk += struct.pack("<Q", corpus.crc64_xz(plain) ^ wrong)
            unpadded += 8
        elif check == 0x0A:
            blk += bytes(32)
            unpadded += 32
        out += blk
        records.append((unpadded, len(plain)))
    if r() < 0.08 and records:
        records[rnd.randrange(len(records))] = (records[0][0] + ch([-4, 4]), records[0][1])  # index disagrees
    if r() < 0.05:
        records = records + [(8, 0)] if r() < 0.5 else records[:-1]
    idx = bytearray(b"\0" + corpus._multibyte(len(records) + (1 if r() < 0.03 else 0)))
    for u, pl in records:
        idx += corpus._multibyte(u) + corpus._multibyte(pl + (1 if r() < 0.03 else 0))
    idx += (b"\0" if r() > 0.04 else b"\2") * ((4 - len(idx) % 4) % 4)
    idx += struct.pack("<I", zlib.crc32(bytes(idx)) ^ (1 if r() < 0.03 else 0))
    fflags = flags if r() > 0.05 else bytes([0, ch([0, 1, 4])])
    fb = struct.pack("<I", (len(idx) // 4 - 1 + (1 if r() < 0.04 else 0)) & 0xFFFFFFFF) + fflags
    out += idx + struct.pack("<I", zlib.crc32(fb) ^ (1 if r() < 0.03 else 0)) + fb + (b"YZ" if r() > 0.03 else b"YY")
    if r() < 0.05:
        out += bytes(ch([1, 4, 12]))
    if r() < 0.05:
        out = out[:rnd.randrange(len(out))]
    return bytes(out)


def _symbols(rnd, enc, n, hist_floor=0):
    """Encodes up to n random symbols with `enc` (tests/corpus.py LzmaEncoder): mostly valid, edge-heavy; now and then a
    distance the window cannot satisfy (the decoder must stop there with the reference's error).  Returns False after
    an invalid symbol (nothing sensible can follow)."""
    lens = [2, 2, 3, 4, 8, 9, 10, 17, 18, 19, 100, 272, 273]
    for _ in range(n):
        h = len(enc.hist) - hist_floor
        k = rnd.random()
        if h == 0 or k < 0.45:
            enc.literal(rnd.randrange(256) if rnd.random() < 0.5 else (enc.hist[-1] if enc.hist else 0))
        elif k < 0.70:
            edges = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 127, 128, 129, 4095, 4096, 4097, 65535, 65536, 65537, h, max(1, h - 1)]
            dist = rnd.choice(edges) if rnd.random() < 0.7 else rnd.randrange(1, h + 1)
            if dist > h:
                dist = rnd.randrange(1, h + 1)
            if rnd.random() < 0.004:
                dist = h + rnd.choice([1, 2, 1000, 1 << 20, 0xFFFFFFFE - h])
            if dist > h:
                enc.match(rnd.choice(lens), dist, check=False)
                return False
            enc.match(rnd.choice(lens), dist)
        elif k < 0.80:
            if enc.rep[0] + 1 > h:  # rep0 points in front of the window: the decoder must refuse it
                enc.rep_match(0, rnd.choice(lens), check=False)
                return False
            enc.shortrep()
        else:
            idx = rnd.randrange(4)
            if enc.rep[idx] + 1 > h:
                enc.rep_match(idx, rnd.choice(lens), check=False)
                return False
            enc.rep_match(idx, rnd.choice(lens))
    return True


def lzma_structured(rnd, corpus):
    """(fmt, stream): a .lzma or raw LZMA2 stream built symbol by symbol with a real range encoder, so that the decoder
    is driven through valid but unusual parses (every properties combination, rep matches before any match, lengths and
    distances at slot / tree boundaries, chunk sequences with every reset mode, size fields that lie)."""
    import struct
    lc, lp, pb = rnd.choice([(3, 0, 2), (0, 0, 0), (4, 0, 4), (0, 4, 0), (2, 2, 1), (1, 3, 3), (8, 4, 4), (5, 0, 2), (3, 2, 0)])
    if rnd.random() < 0.5:  # ---- .lzma
        enc = corpus.LzmaEncoder(lc, lp, pb)
        ok = _symbols(rnd, enc, rnd.choice([0, 1, 5, 40, 300, 2500]))
        marker = rnd.random() < 0.6
        if ok and marker:
            enc.end_marker()
        n = len(enc.hist)
        size = None if (marker and rnd.random() < 0.7) else max(0, n + rnd.choice([0, 0, 0, -1, 1, -n // 2, 300]))
        body = corpus.lzma_header(lc, lp, pb, rnd.choice([0, 4096, 1 << 16, 1 << 20, 0x7F7F7F7F]), size) + enc.finish()
        if rnd.random() < 0.1:
            body += bytes(rnd.choice([1, 5, 20]))
        if rnd.random() < 0.1 and len(body) > 14:
            body = body[:rnd.randrange(13, len(body))]
        return 0, body
    # ---- LZMA2
    if lc + lp > 4 and rnd.random() < 0.8:
        lc, lp, pb = 3, 0, 2
    enc = corpus.LzmaEncoder(lc, lp, pb)
    out = bytearray()
    floor = 0            # first byte of the current dictionary epoch inside enc.hist
    have_props = False
    for ci in range(rnd.choice([1, 1, 2, 3, 5])):
        if rnd.random() < 0.25:  # stored chunk
            st = rnd.choice([1, 2]) if ci else rnd.choice([1, 1, 2])
            data = bytes(rnd.randrange(256) for _ in range(rnd.choice([1, 2, 50, 700])))
            out += bytes([st]) + struct.pack(">H", len(data) - 1) + data
            if st == 1:
                floor = len(enc.hist)
            enc.hist += data
            continue
        ctrl = rnd.choice([0xE0, 0xC0, 0xA0, 0x80]) if have_props else rnd.choice([0xE0, 0xE0, 0xE0, 0xC0, 0xA0, 0x80])
        props = None
        if ctrl >= 0xC0:
            if rnd.random() < 0.3:
                lc, lp, pb = rnd.choice([(3, 0, 2), (0, 0, 0), (4, 0, 4), (0, 4, 0), (2, 2, 1), (1, 3, 3), (3, 2, 0)])
            props = corpus.props_byte(lc, lp, pb)
            have_props = True
        elif not have_props and ctrl < 0xC0:
            lc, lp, pb = 0, 0, 0  # the reference decodes with lc=lp=pb=0 until the first props byte (lzma2.rs:23-34)
        if ctrl == 0xE0:
            floor = len(enc.hist)
        if ctrl >= 0xA0:
            enc.reset_state(lc, lp, pb)
        enc.new_chunk()
        before = len(enc.hist)
        ok = _symbols(rnd, enc, rnd.choice([1, 3, 30, 200, 1500]), hist_floor=floor)
        if ok and len(enc.hist) == before:
            enc.literal(rnd.randrange(256))
        if ok and rnd.random() < 0.03:
            enc.end_marker()  # an end marker inside an LZMA2 chunk
        payload = enc.finish()
        unpacked = len(enc.hist) - before
        claim = max(1, unpacked + (0 if rnd.random() > 0.08 else rnd.choice([-1, 1, 5])))
        packed = max(1, len(payload) + (0 if rnd.random() > 0.08 else rnd.choice([-1, 1, 3])))
        hdr = bytes([ctrl | ((claim - 1) >> 16)]) + struct.pack(">HH", (claim - 1) & 0xFFFF, (packed - 1) & 0xFFFF)
        if props is not None:
            hdr += bytes([props if rnd.random() > 0.03 else rnd.choice([225, 255, corpus.props_byte(4, 1, 0)])])
        out += hdr + payload
        if not ok:
            break
    out += b"\0" if rnd.random() > 0.1 else bytes([rnd.choice([3, 0x7F])])
    if rnd.random() < 0.1:
        out += b"trailing"
    if rnd.random() < 0.08 and len(out) > 2:
        out = out[:rnd.randrange(1, len(out))]
    return 1, bytes(out)


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
    ap.add_argument("--backend", choices=["emul", "gpu", "gpu-device"], default="emul")
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--per-format", type=int, default=400)
    ap.add_argument("--lzma-structured", action="store_true",
                    help=".lzma / LZMA2 streams built symbol by symbol with a real range encoder (state-machine fuzz)")
    ap.add_argument("--xz-structured", action="store_true",
                    help=".xz files assembled field by field with valid CRCs around odd values (container walk fuzz)")
    ap.add_argument("--cases", action="store_true",
                    help="mutate the streams of the parity corpus (tests/cases.py: hand-encoded, chained .xz, ...) instead")
    a = ap.parse_args()
    import corpus
    import parity
    if a.backend == "gpu-device":
        # the device-resident entry point (K2 scan on the device, capacities fixed by the caller, no retry): .lzma and raw
        # LZMA2 only; a stream that reports LZB_E_CAPACITY is decoded again through the host path, which retries
        import gpu_util
        from lzma_rs_b200 import Context
        ctx = Context()

        def decode(fmt, streams, opts):
            if fmt == 2:
                return gpu_util.host_decode(ctx, fmt, streams, opts)
            caps = ctx.scan(fmt, *__import__("lzma_rs_b200")._native.pack_streams(streams), gpu_util.options_from(opts))
            b = gpu_util.DeviceBatch(ctx, fmt, streams, caps, opts).decode()
            res = []
            for i, s in enumerate(streams):
                if int(b.st[i]["code"]) == -1:
                    res.append(gpu_util.host_decode(ctx, fmt, [s], opts)[0])
                else:
                    res.append(test_emul_parity.emul_py.Result(b.output(i), int(b.consumed[i]), b.st[i].copy(), b.display(i)))
            return res
        import test_emul_parity
    elif a.backend == "gpu":
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
    for rd in range(a.rounds if a.lzma_structured else 0):
        gen = [lzma_structured(rnd, corpus) for _ in range(a.per_format)]
        # .lzma streams also run under the other decompress::Options (options.rs:24-43): sizes provided from outside,
        # small memory limits -- the same streams, the same options on both sides
        opt_sets = [{}, {"unpacked_mode": 1, "provided": rnd.choice([0, 1, 7, 300])}, {"unpacked_mode": 1},
                    {"memlimit": rnd.choice([0, 1, 100, 4096, 70_000])}, {"unpacked_mode": 2, "provided": rnd.choice([0, 5, 1000])},
                    {"unpacked_mode": 2}]
        for fmt, opts in [(0, o) for o in ([{}] + rnd.sample(opt_sets[1:], 2))] + [(1, {})]:
            named = [(f"r{rd}-f{fmt}-{i}", st) for i, (f, st) in enumerate(gen) if f == fmt]
            bad = parity.check_group(decode, fmt, opts, named)
            if bad:
                print(f"   options: {opts}", flush=True)
            total += len(named)
            if bad:
                bad_total += len(bad)
                print(f"round {rd} structured fmt {fmt}: {len(bad)} mismatches", flush=True)
                for line in bad[:5]:
                    print("   " + line, flush=True)
                names = {b.split(":")[0] for b in bad}
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                for name, st in named:
                    if name in names:
                        open(os.path.join(ROOT, "gpurun_out", f"fuzzfail-{a.backend}-l{a.seed}-{name}.bin"), "wb").write(st)
    for rd in range(a.rounds if a.xz_structured else 0):
        named = [(f"r{rd}-f2-{i}", xz_structured(rnd, corpus)) for i in range(a.per_format)]
        bad = parity.check_group(decode, 2, {}, named)
        total += len(named)
        if bad:
            bad_total += len(bad)
            print(f"round {rd} structured xz: {len(bad)} mismatches", flush=True)
            for line in bad[:5]:
                print("   " + line, flush=True)
            names = {b.split(":")[0] for b in bad}
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            for name, st in named:
                if name in names:
                    open(os.path.join(ROOT, "gpurun_out", f"fuzzfail-{a.backend}-x{a.seed}-{name}.bin"), "wb").write(st)
    for rd in range(0 if (a.xz_structured or a.lzma_structured) else a.rounds):
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
