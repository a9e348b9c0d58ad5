"""Parity cases shared by the CPU-tier emulation tests and the GPU tests.

Each case = (name, fmt, stream bytes, options kwargs).  fmt: 0 LZMA, 1 LZMA2, 2 XZ.  Expected results always
come from the oracle (tests/oracle_py.py) at test time; valid streams are additionally cross-checked against
liblzma where liblzma accepts them (the reference's own differential oracle, tests/lzma.rs:109-114).
"""
import random
import struct
import zlib

import numpy as np

import corpus as c

LZMA, LZMA2, XZ = 0, 1, 2


def _zeros(n):
    return b"\0" * n


def _rand(seed, n):
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8).tobytes()


def valid_lzma2_cases():
    out = []
    datas = [("empty", b""), ("one", b"x"), ("two", b"ab"), ("text100", c.mixed_text(1, 100)),
             ("text64k", c.mixed_text(2, 65536)), ("text300k", c.mixed_text(3, 300_000)),
             ("zeros1m", _zeros(1_000_000)), ("ff100k", b"\xff" * 100_000), ("rand70k", _rand(4, 70_000)),
             ("randtext", _rand(5, 3000) + c.mixed_text(6, 50_000) + _rand(7, 70_000) + c.mixed_text(8, 9000))]
    for name, d in datas:
        out.append((f"l2-{name}", LZMA2, c.raw_lzma2(d, dict_size=1 << 20), {}))
    # property sweep incl. lc+lp = 4, pb 0..4, lp != 0
    d = c.mixed_text(9, 40_000)
    for lc, lp, pb in [(0, 0, 0), (3, 0, 2), (4, 0, 4), (0, 4, 0), (2, 2, 1), (1, 3, 3), (3, 1, 2), (0, 2, 4), (2, 1, 0)]:
        out.append((f"l2-props-{lc}{lp}{pb}", LZMA2, c.raw_lzma2(d, dict_size=1 << 16, lc=lc, lp=lp, pb=pb), {}))
    # small dictionary -> more chunks with dict resets / far matches clipped
    out.append(("l2-dict4k", LZMA2, c.raw_lzma2(c.mixed_text(10, 200_000), dict_size=4096), {}))
    out.append(("l2-preset0", LZMA2, c.raw_lzma2(c.mixed_text(11, 100_000), preset=0), {}))
    out.append(("l2-preset9", LZMA2, c.raw_lzma2(c.mixed_text(12, 100_000), preset=9, dict_size=1 << 20), {}))
    # the reference's own encoder shape: stored chunks only (src/encode/lzma2.rs)
    for name, d in [("empty", b""), ("hello", b"Hello world"), ("zeros", _zeros(200_000)), ("text", c.mixed_text(13, 70_000))]:
        out.append((f"l2-stored-{name}", LZMA2, c.stored_lzma2(d), {}))
    out.append(("l2-stress-rep0", LZMA2, c.rep0_stress_lzma2(262144), {}))
    out.append(("l2-stress-rep0-small", LZMA2, c.rep0_stress_lzma2(1000, byte=0), {}))
    # trailing bytes after the end of the stream stay unread (consumed < len)
    out.append(("l2-trailing", LZMA2, c.raw_lzma2(c.mixed_text(14, 5000)) + b"TRAILING", {}))
    return out


def valid_lzma_cases():
    out = []
    for name, d in [("empty", b""), ("one", b"x"), ("text64k", c.mixed_text(20, 65536)), ("zeros", _zeros(300_000)),
                    ("rand", _rand(21, 20_000)), ("text200k", c.mixed_text(22, 200_000))]:
        out.append((f"lz-{name}", LZMA, c.lzma_alone(d, dict_size=1 << 16), {}))
        out.append((f"lz-known-{name}", LZMA, c.lzma_alone_known_size(d, dict_size=1 << 16), {}))
    d = c.mixed_text(23, 50_000)
    out.append(("lz-dict4k-wrap", LZMA, c.lzma_alone(d, dict_size=4096), {}))  # ring wraps many times
    out.append(("lz-dict-small-header", LZMA, c.lzma_alone(d, dict_size=4096)[:1] + struct.pack("<I", 16) +
                c.lzma_alone(d, dict_size=4096)[5:], {}))  # header dict < 0x1000 is raised to 0x1000 (lzma.rs:122-126)
    for lc, lp, pb in [(0, 0, 0), (4, 0, 2), (0, 4, 4), (2, 2, 2), (1, 1, 1)]:
        out.append((f"lz-props-{lc}{lp}{pb}", LZMA, c.lzma_alone(d, lc=lc, lp=lp, pb=pb), {}))
    # reference's own dumb encoder shapes (src/encode/dumbencoder.rs) + every UnpackedSize mode (tests/lzma.rs:237-303)
    data = b"Some data"
    out.append(("lz-dumb-marker", LZMA, c.dumb_lzma(data), {}))
    out.append(("lz-dumb-size", LZMA, c.dumb_lzma(data, unpacked_in_header=len(data)), {}))
    out.append(("lz-dumb-nosizefield", LZMA, c.dumb_lzma(data, write_size_field=False), {"unpacked_mode": 2, "provided": len(data)}))
    out.append(("lz-dumb-some-useprov", LZMA, c.dumb_lzma(data, unpacked_in_header=len(data)), {"unpacked_mode": 1, "provided": len(data)}))
    out.append(("lz-dumb-none-useprov", LZMA, c.dumb_lzma(data), {"unpacked_mode": 1, "provided": len(data)}))
    out.append(("lz-dumb-none-provnone", LZMA, c.dumb_lzma(data), {"unpacked_mode": 1, "provided": None}))
    out.append(("lz-dumb-1m-zeros", LZMA, c.dumb_lzma(_zeros(100_000)), {}))
    # memlimit (tests/lzma.rs:306-336 and beyond)
    out.append(("lz-memlimit0", LZMA, c.dumb_lzma(data), {"unpacked_mode": 1, "provided": None, "memlimit": 0}))
    big = c.lzma_alone(c.mixed_text(24, 30_000), dict_size=1 << 16)
    for ml in (1, 100, 4095, 4096, 29_999, 30_000, 65_535, 65_536, 1 << 30):
        out.append((f"lz-memlimit-{ml}", LZMA, big, {"memlimit": ml}))
    out.append(("lz-memlimit-dict4k", LZMA, c.lzma_alone(c.mixed_text(24, 30_000), dict_size=4096), {"memlimit": 5000}))
    # provided size smaller / larger than the real one
    out.append(("lz-prov-small", LZMA, big, {"unpacked_mode": 1, "provided": 1000}))
    out.append(("lz-prov-large", LZMA, big, {"unpacked_mode": 1, "provided": 40_000}))
    out.append(("lz-trailing-known", LZMA, c.lzma_alone_known_size(d) + b"XYZ", {}))
    out.append(("lz-trailing-marker", LZMA, c.lzma_alone(d) + b"XYZ", {}))  # marker then extra bytes -> error
    return out


def _clone_model(enc):
    """A new encoder (fresh range coder) whose probabilities / state / reps continue from `enc` (LZMA2 chunk 0x80)."""
    import copy
    e = copy.deepcopy(enc)
    e.new_chunk()
    return e


def hand_encoded_cases():
    """Streams built symbol by symbol: every symbol kind, edge lengths/distances, and malformed constructions."""
    out = []
    rnd = random.Random(1234)
    for t in range(6):
        lc, lp, pb = [(3, 0, 2), (0, 0, 0), (4, 0, 4), (0, 4, 1), (2, 2, 3), (1, 0, 2)][t]
        enc = c.LzmaEncoder(lc, lp, pb)
        for _ in range(1500):
            k = rnd.random()
            n = len(enc.hist)
            if n < 4 or k < 0.35:
                enc.literal(rnd.randrange(256) if rnd.random() < 0.5 else 65)
            elif k < 0.65:
                enc.match(rnd.choice([2, 3, 4, 9, 10, 17, 18, 273, rnd.randrange(2, 274)]),
                          rnd.choice([1, 2, n, max(1, n - 1), rnd.randrange(1, n + 1)]))
            elif k < 0.78:
                enc.shortrep() if enc.rep[0] + 1 <= n else enc.literal(1)
            else:
                idx = rnd.randrange(4)
                if enc.rep[idx] + 1 <= n:
                    enc.rep_match(idx, rnd.choice([2, 8, 9, 16, 17, 272, 273, rnd.randrange(2, 274)]))
                else:
                    enc.literal(2)
        plain = bytes(enc.hist)
        enc.end_marker()
        payload = enc.finish()
        out.append((f"hand-lzma-{t}", LZMA, c.lzma_header(lc, lp, pb, 1 << 20) + payload, {}))
        out.append((f"hand-lzma-known-{t}", LZMA, c.lzma_header(lc, lp, pb, 1 << 20, len(plain)) + payload, {}))
        out.append((f"hand-lzma-dict4k-{t}", LZMA, c.lzma_header(lc, lp, pb, 4096) + payload, {}))  # far dists -> error
        # declared size smaller (overshoot by a match likely) / larger (marker hit first)
        out.append((f"hand-lzma-short-{t}", LZMA, c.lzma_header(lc, lp, pb, 1 << 20, len(plain) // 2) + payload, {}))
        out.append((f"hand-lzma-long-{t}", LZMA, c.lzma_header(lc, lp, pb, 1 << 20, len(plain) + 7) + payload, {}))

    # --- LZMA2 constructions
    P = c.props_byte(3, 0, 2)

    def chunk_stream(build, n_lit=40):
        enc = c.LzmaEncoder(3, 0, 2)
        build(enc)
        return enc

    # (a) chunk A ends with a match (state >= 7); stored chunk with dict reset; chunk 0x80 starts with a literal
    #     -> matched-literal lookup beyond the fresh window ("Match distance .. beyond output size")
    enc = c.LzmaEncoder(3, 0, 2)
    for b in b"abcdefgh":
        enc.literal(b)
    enc.match(5, 8)
    a_plain = len(enc.hist)
    a = c.lzma2_chunk(enc.finish(), a_plain, 0xE0, P)
    stored = bytes([1]) + struct.pack(">H", 2) + b"XYZ"
    enc2 = _clone_model(enc)
    enc2.hist = bytearray(b"0123456789a")  # pretend window (len & 3 == 3 like the decoder's) so literal() can run
    enc2.literal(0x55)
    b_chunk = c.lzma2_chunk(enc2.finish(), 1, 0x80)
    out.append(("hand-l2-matchbyte-after-dictreset", LZMA2, a + stored + b_chunk + b"\0", {}))
    # same but stored chunk WITHOUT dict reset (status 2): matched literal reads into the stored bytes
    stored2 = bytes([2]) + struct.pack(">H", 2) + b"XYZ"
    enc3 = _clone_model(enc)
    enc3.hist = bytearray(enc.hist) + bytearray(b"XYZ")
    enc3.literal(0x55)
    enc3.shortrep()
    enc3.rep_match(0, 20)
    out.append(("hand-l2-matchbyte-after-stored", LZMA2, a + stored2 + c.lzma2_chunk(enc3.finish(), 22, 0x80) + b"\0", {}))

    # (b) first chunk without any reset (0x80): decodes with lc=lp=pb=0 (reference leniency, lzma2.rs:23-34)
    enc = c.LzmaEncoder(0, 0, 0)
    for b in b"no reset at all, lc=lp=pb=0":
        enc.literal(b)
    enc.match(6, 3)
    n = len(enc.hist)
    out.append(("hand-l2-first-chunk-no-reset", LZMA2, c.lzma2_chunk(enc.finish(), n, 0x80) + b"\0", {}))
    # state reset without props (0xA0) as first chunk
    enc = c.LzmaEncoder(0, 0, 0)
    for b in b"state reset only":
        enc.literal(b)
    n = len(enc.hist)
    out.append(("hand-l2-first-chunk-state-reset", LZMA2, c.lzma2_chunk(enc.finish(), n, 0xA0) + b"\0", {}))

    # (c) chunk whose packed size is larger than what the range decoder needs: the reference does NOT skip the
    #     remainder (lzma2.rs:189-192) -> the leftover bytes are parsed as control bytes
    enc = c.LzmaEncoder(3, 0, 2)
    for b in b"under-consumed chunk":
        enc.literal(b)
    n = len(enc.hist)
    payload = enc.finish()
    for extra, nm in [(b"\0", "zero"), (b"\x01\x00\x02abc\x00", "stored"), (b"\x05", "invalid"), (b"\0\0\0\0", "zeros")]:
        s = c.lzma2_chunk(payload + extra, n, 0xE0, P) + b"\0"
        out.append((f"hand-l2-underconsume-{nm}", LZMA2, s, {}))

    # (d) end marker inside an LZMA2 chunk; overshooting match; unpacked size too large for the data
    enc = c.LzmaEncoder(3, 0, 2)
    for b in b"marker inside":
        enc.literal(b)
    n = len(enc.hist)
    enc.end_marker()
    out.append(("hand-l2-endmarker", LZMA2, c.lzma2_chunk(enc.finish(), n + 5, 0xE0, P) + b"\0", {}))
    enc = c.LzmaEncoder(3, 0, 2)
    for b in b"overshoot":
        enc.literal(b)
    enc.match(100, 3)
    out.append(("hand-l2-overshoot", LZMA2, c.lzma2_chunk(enc.finish(), 50, 0xE0, P) + b"\0", {}))
    enc = c.LzmaEncoder(3, 0, 2)
    for b in b"short data":
        enc.literal(b)
    out.append(("hand-l2-unpacked-too-large", LZMA2, c.lzma2_chunk(enc.finish(), 500, 0xE0, P) + b"\0", {}))

    # (e) distance beyond the window (match and rep), in LZMA2 and LZMA
    enc = c.LzmaEncoder(3, 0, 2)
    for b in b"abc":
        enc.literal(b)
    enc.match(4, 10, check=False)
    out.append(("hand-l2-dist-beyond", LZMA2, c.lzma2_chunk(enc.finish(), 7, 0xE0, P) + b"\0", {}))
    enc = c.LzmaEncoder(3, 0, 2)
    for b in b"abc":
        enc.literal(b)
    enc.match(4, 10, check=False)
    out.append(("hand-lzma-dist-beyond", LZMA, c.lzma_header(3, 0, 2, 1 << 16, 7) + enc.finish(), {}))
    enc = c.LzmaEncoder(3, 0, 2)
    for i in range(5000):
        enc.literal(i & 0xFF)
    enc.match(4, 4500)
    enc.end_marker()
    out.append(("hand-lzma-dist-beyond-dict", LZMA, c.lzma_header(3, 0, 2, 4096) + enc.finish(), {}))

    # (f) LZMA2 header errors
    good = c.raw_lzma2(c.mixed_text(30, 3000))
    out.append(("hand-l2-invalid-status-3", LZMA2, b"\x03" + good, {}))
    out.append(("hand-l2-invalid-status-7f", LZMA2, good[:-1] + b"\x7f", {}))
    out.append(("hand-l2-props-225", LZMA2, good[:5] + bytes([225]) + good[6:], {}))
    out.append(("hand-l2-props-lclp5", LZMA2, good[:5] + bytes([c.props_byte(3, 2, 0)]) + good[6:], {}))
    out.append(("hand-l2-props-lc8", LZMA2, good[:5] + bytes([c.props_byte(8, 0, 0)]) + good[6:], {}))
    out.append(("hand-l2-no-end", LZMA2, good[:-1], {}))
    out.append(("hand-l2-stored-short", LZMA2, b"\x01\x00\x09abc", {}))
    out.append(("hand-l2-stored-nosize", LZMA2, b"\x01\x00", {}))
    out.append(("hand-l2-stored-status2-first", LZMA2, b"\x02\x00\x02abc\x00", {}))  # no dict reset at all: allowed
    out.append(("hand-l2-empty-input", LZMA2, b"", {}))
    out.append(("hand-l2-just-zero", LZMA2, b"\0", {}))
    out.append(("hand-l2-chunk-short-hdr", LZMA2, b"\xe0\x00", {}))
    out.append(("hand-l2-chunk-short-hdr2", LZMA2, b"\xe0\x00\x10\x00", {}))
    out.append(("hand-l2-chunk-no-props", LZMA2, b"\xe0\x00\x10\x00\x10", {}))
    out.append(("hand-l2-chunk-short-rc", LZMA2, b"\xe0\x00\x10\x00\x10\x5d\x00\x01", {}))
    return out


def truncation_and_corruption_cases():
    out = []
    small2 = c.raw_lzma2(c.mixed_text(40, 600))
    for cut in range(len(small2)):
        out.append((f"trunc-l2-{cut}", LZMA2, small2[:cut], {}))
    small1 = c.lzma_alone(c.mixed_text(41, 400), dict_size=4096)
    for cut in range(len(small1)):
        out.append((f"trunc-lz-{cut}", LZMA, small1[:cut], {}))
    small1k = c.lzma_alone_known_size(c.mixed_text(41, 400), dict_size=4096)
    for cut in range(13, len(small1k), 7):
        out.append((f"trunc-lzk-{cut}", LZMA, small1k[:cut], {}))
    rnd = random.Random(99)
    base2 = c.raw_lzma2(c.mixed_text(42, 20_000))
    base1 = c.lzma_alone(c.mixed_text(43, 20_000), dict_size=1 << 14)
    for t in range(120):
        b = bytearray(base2)
        for _ in range(rnd.choice([1, 1, 2, 5])):
            b[rnd.randrange(len(b))] = rnd.randrange(256)
        out.append((f"corrupt-l2-{t}", LZMA2, bytes(b), {}))
        b = bytearray(base1)
        for _ in range(rnd.choice([1, 1, 2, 5])):
            b[rnd.randrange(13 if rnd.random() < 0.1 else len(b))] = rnd.randrange(256)
        out.append((f"corrupt-lz-{t}", LZMA, bytes(b), {}))
    for t in range(40):
        out.append((f"garbage-l2-{t}", LZMA2, bytes([rnd.choice([0x80, 0xE0, 0xC0, 0xA0, 1, 2])]) + _rand(1000 + t, rnd.randrange(1, 300)), {}))
        g = bytearray(_rand(2000 + t, rnd.randrange(14, 300)))
        g[0] = rnd.randrange(225)
        if t % 2:
            g[5:13] = b"\xff" * 8
        out.append((f"garbage-lz-{t}", LZMA, bytes(g), {}))
    out.append(("garbage-lz-text", LZMA, b"corrupted bytes here corrupted bytes here", {}))  # stream.rs:461-467
    out.append(("garbage-lz-ff", LZMA, b"\xff" * 32, {}))  # stream.rs:376-388
    return out


def xz_cases():
    out = []
    big = c.mixed_text(50, 700_000)
    out.append(("xz-multi-crc32", XZ, c.xz_file(big, block_size=1 << 18, check=c.CHECK_CRC32), {}))
    out.append(("xz-multi-crc64-sizes", XZ, c.xz_file(big, block_size=100_000, check=c.CHECK_CRC64, with_sizes=True), {}))
    out.append(("xz-multi-none", XZ, c.xz_file(big[:200_000], block_size=50_000, check=c.CHECK_NONE), {}))
    out.append(("xz-sha256", XZ, c.xz_file(big[:10_000], check=c.CHECK_SHA256), {}))
    out.append(("xz-empty", XZ, c.xz_file(b""), {}))
    out.append(("xz-one-byte", XZ, c.xz_file(b"x", check=c.CHECK_CRC64), {}))
    out.append(("xz-many-small-blocks", XZ, c.xz_file(c.mixed_text(51, 20_000), block_size=512, check=c.CHECK_CRC32), {}))
    import lzma as _l
    out.append(("xz-liblzma-default", XZ, _l.compress(big[:300_000]), {}))  # CRC64, sizes absent, one block
    out.append(("xz-liblzma-crc32", XZ, _l.compress(big[:100_000], check=_l.CHECK_CRC32), {}))
    out.append(("xz-liblzma-none", XZ, _l.compress(big[:100_000], check=_l.CHECK_NONE), {}))
    out.append(("xz-liblzma-sha256", XZ, _l.compress(big[:100_000], check=_l.CHECK_SHA256), {}))
    # two concatenated streams / stream padding are rejected (xz.rs:88-92)
    one = c.xz_file(b"hello world", check=c.CHECK_CRC32)
    out.append(("xz-concatenated", XZ, one + one, {}))
    out.append(("xz-padding", XZ, one + b"\0\0\0\0", {}))

    small = c.xz_file(c.mixed_text(52, 3000), block_size=1000, check=c.CHECK_CRC32, with_sizes=True)
    for cut in range(0, len(small), 3):
        out.append((f"xz-trunc-{cut}", XZ, small[:cut], {}))
    small64 = c.xz_file(c.mixed_text(53, 2500), block_size=1200, check=c.CHECK_CRC64)
    for pos in range(len(small64)):  # flip every byte of a small 3-block file once
        b = bytearray(small64)
        b[pos] ^= 0x5A
        out.append((f"xz-flip64-{pos}", XZ, bytes(b), {}))
    for pos in range(0, len(small), 2):
        b = bytearray(small)
        b[pos] = (b[pos] + 1) & 0xFF
        out.append((f"xz-inc-{pos}", XZ, bytes(b), {}))

    # hand-built container errors with consistent CRCs so the deeper checks are reached
    plain = c.mixed_text(54, 5000)
    payload = c.raw_lzma2(plain)

    def container(blocks_bytes, records, check=c.CHECK_CRC32, footer_check=None, index_tweak=None, backward_tweak=0,
                  footer_magic=b"YZ"):
        flags = bytes([0, check])
        o = bytearray(b"\xfd7zXZ\0" + flags + struct.pack("<I", zlib.crc32(flags)))
        o += blocks_bytes
        idx = bytearray(b"\0" + c._multibyte(len(records) if index_tweak != "count" else len(records) + 1))
        for u, p in records:
            if index_tweak == "unpadded":
                u += 4
            if index_tweak == "unpacked":
                p += 1
            idx += c._multibyte(u) + c._multibyte(p)
        pad = (4 - len(idx) % 4) % 4
        idx += (b"\0" if index_tweak != "padding" else b"\1") * pad
        crc = zlib.crc32(bytes(idx))
        if index_tweak == "crc":
            crc ^= 1
        idx += struct.pack("<I", crc)
        o += idx
        fflags = bytes([0, check if footer_check is None else footer_check])
        fb = struct.pack("<I", len(idx) // 4 - 1 + backward_tweak) + fflags
        o += struct.pack("<I", zlib.crc32(fb)) + fb + footer_magic
        return bytes(o)

    blk, unp = c.xz_block(payload, plain, c.CHECK_CRC32)
    rec = [(unp, len(plain))]
    out.append(("xz-hand-ok", XZ, container(blk, rec), {}))
    for tweak in ("count", "unpadded", "unpacked", "padding", "crc"):
        out.append((f"xz-hand-index-{tweak}", XZ, container(blk, rec, index_tweak=tweak), {}))
    out.append(("xz-hand-backward", XZ, container(blk, rec, backward_tweak=1), {}))
    out.append(("xz-hand-footer-flags", XZ, container(blk, rec, footer_check=c.CHECK_CRC64), {}))
    out.append(("xz-hand-footer-badcheck", XZ, container(blk, rec, footer_check=0x02), {}))
    out.append(("xz-hand-footer-magic", XZ, container(blk, rec, footer_magic=b"ZY"), {}))
    out.append(("xz-hand-trailing", XZ, container(blk, rec) + b"x", {}))
    # block header variants
    def hdr_block(body_fn, payload=payload, plain=plain, check=c.CHECK_CRC32):
        body = body_fn()
        total = 1 + len(body) + 4
        total_padded = (total + 3) & ~3
        body += b"\0" * (total_padded - total)
        hdr = bytes([total_padded // 4 - 1]) + body
        hdr += struct.pack("<I", zlib.crc32(hdr))
        blk = hdr + payload
        unp = len(blk)
        blk += b"\0" * ((4 - len(blk) % 4) % 4)
        blk += struct.pack("<I", zlib.crc32(plain))
        return blk, unp + 4

    variants = {
        "reserved-flags": lambda: bytes([0x04, 0x21, 0x01, 0x16]),
        "unknown-filter": lambda: bytes([0x00, 0x03, 0x01, 0x16]),
        "props-too-big": lambda: bytes([0x00, 0x21, 0x7f, 0x16]),
        "props-len2": lambda: bytes([0x00, 0x21, 0x02, 0x16, 0x00]),
        "props-len0": lambda: bytes([0x00, 0x21, 0x00]),
        "wrong-packed": lambda: bytes([0x40]) + c._multibyte(len(payload) + 1) + bytes([0x21, 0x01, 0x16]),
        "wrong-unpacked": lambda: bytes([0x80]) + c._multibyte(len(plain) - 1) + bytes([0x21, 0x01, 0x16]),
        "right-sizes": lambda: bytes([0xC0]) + c._multibyte(len(payload)) + c._multibyte(len(plain)) + bytes([0x21, 0x01, 0x16]),
        "multibyte-overlong": lambda: bytes([0x40]) + b"\xff" * 9 + b"\x01" + bytes([0x21, 0x01, 0x16]),
        "hdr-padding-nonzero": lambda: bytes([0x00, 0x21, 0x01, 0x16, 0x01]),
    }
    for nm, fn in variants.items():
        blk2, unp2 = hdr_block(fn)
        out.append((f"xz-hand-hdr-{nm}", XZ, container(blk2, [(unp2, len(plain))]), {}))
    # bad block padding / bad block CRC with everything else consistent
    b3 = bytearray(blk)
    if (len(payload) + 12) % 4:
        b3[12 + len(payload)] = 1
        out.append(("xz-hand-block-padding", XZ, container(bytes(b3), rec), {}))
    b4 = bytearray(blk)
    b4[-1] ^= 0xFF
    out.append(("xz-hand-block-crc", XZ, container(bytes(b4), rec), {}))
    blk64, unp64 = c.xz_block(payload, plain, c.CHECK_CRC64)
    b5 = bytearray(blk64)
    b5[-3] ^= 0x10
    out.append(("xz-hand-block-crc64", XZ, container(bytes(b5), [(unp64, len(plain))], check=c.CHECK_CRC64), {}))
    # LZMA2 payload that ends early / is malformed inside a container; second block after a bad first block
    bad_payload = payload[:-1] + b"\x7f"
    blkb, unpb = c.xz_block(bad_payload, plain, c.CHECK_CRC32)
    out.append(("xz-hand-bad-lzma2", XZ, container(blkb, [(unpb, len(plain))]), {}))
    out.append(("xz-hand-good-then-bad", XZ, container(blk + blkb, rec + [(unpb, len(plain))]), {}))
    # under-consuming LZMA2 chunk inside a block: framing scan mispredicts, look-ahead must be redone
    enc = c.LzmaEncoder(3, 0, 2)
    for b in b"under-consumed chunk in xz":
        enc.literal(b)
    n = len(enc.hist)
    upayload = c.lzma2_chunk(enc.finish() + b"\0", n, 0xE0, c.props_byte(3, 0, 2)) + b"\0"
    uplain = bytes(enc.hist)
    # the decoder stops at the first 0x00 (inside the chunk's slack); the second 0x00 becomes block padding / garbage
    blku, unpu = c.xz_block(upayload[:-1], uplain, c.CHECK_CRC32)
    out.append(("xz-hand-underconsume", XZ, container(blku + blk, [(unpu, len(uplain))] + rec), {}))
    return out


def _xz_wrap(blocks_bytes, records, check):
    flags = bytes([0, check])
    o = bytearray(b"\xfd7zXZ\0" + flags + struct.pack("<I", zlib.crc32(flags))) + blocks_bytes
    idx = bytearray(b"\0" + c._multibyte(len(records)))
    for u, p in records:
        idx += c._multibyte(u) + c._multibyte(p)
    idx += b"\0" * ((4 - len(idx) % 4) % 4)
    idx += struct.pack("<I", zlib.crc32(bytes(idx)))
    fb = struct.pack("<I", len(idx) // 4 - 1) + flags
    return bytes(o + idx + struct.pack("<I", zlib.crc32(fb)) + fb + b"YZ")


def xz_chain_cases():
    """Blocks with 2-4 chained LZMA2 filters (the reference decodes filter i+1 from filter i's output, xz.rs:240-249;
    xz itself never writes such files)."""
    out = []
    plain = c.mixed_text(60, 30_000)
    lvl = [plain]
    for k in range(4):
        lvl.append(c.raw_lzma2(lvl[-1], dict_size=1 << 16))
    for nf in (2, 3, 4):
        for check in (c.CHECK_CRC32, c.CHECK_CRC64, c.CHECK_NONE):
            blk, unp = c.xz_block(lvl[nf], plain, check, nfilters=nf)
            out.append((f"xz-chain-{nf}-{check}", XZ, _xz_wrap(blk, [(unp, len(plain))], check), {}))
    # a normal block, then a chained one, then a normal one (look-ahead must stop and resume around the chain)
    p2 = c.mixed_text(61, 5000)
    b1, u1 = c.xz_block(c.raw_lzma2(p2), p2, c.CHECK_CRC32)
    b2, u2 = c.xz_block(lvl[2], plain, c.CHECK_CRC32, nfilters=2)
    out.append(("xz-chain-mixed", XZ, _xz_wrap(b1 + b2 + b1, [(u1, len(p2)), (u2, len(plain)), (u1, len(p2))], c.CHECK_CRC32), {}))
    # sizes in the header: packed = filter 0's input, unpacked = the last filter's output
    blk, unp = c.xz_block(lvl[2], plain, c.CHECK_CRC32, with_sizes=True, nfilters=2)
    out.append(("xz-chain-sizes", XZ, _xz_wrap(blk, [(unp, len(plain))], c.CHECK_CRC32), {}))
    # inner stream corrupt / truncated / with trailing bytes after its end marker (ignored by the reference)
    bad_inner = c.raw_lzma2(lvl[1][:-1] + b"\x7f", dict_size=1 << 16)
    blk, unp = c.xz_block(bad_inner, plain, c.CHECK_CRC32, nfilters=2)
    out.append(("xz-chain-inner-bad-status", XZ, _xz_wrap(blk, [(unp, len(plain))], c.CHECK_CRC32), {}))
    trunc_inner = c.raw_lzma2(lvl[1][:len(lvl[1]) // 2], dict_size=1 << 16)
    blk, unp = c.xz_block(trunc_inner, plain, c.CHECK_CRC32, nfilters=2)
    out.append(("xz-chain-inner-truncated", XZ, _xz_wrap(blk, [(unp, len(plain))], c.CHECK_CRC32), {}))
    trail_inner = c.raw_lzma2(lvl[1] + b"trailing garbage", dict_size=1 << 16)
    blk, unp = c.xz_block(trail_inner, plain, c.CHECK_CRC32, nfilters=2)
    out.append(("xz-chain-inner-trailing", XZ, _xz_wrap(blk, [(unp, len(plain))], c.CHECK_CRC32), {}))
    # wrong check over the final output; second filter with 2-byte properties (fails when that filter is reached)
    blk, unp = c.xz_block(lvl[2], plain[:-1] + b"!", c.CHECK_CRC32, nfilters=2)
    out.append(("xz-chain-bad-crc", XZ, _xz_wrap(blk, [(unp, len(plain))], c.CHECK_CRC32), {}))
    body = bytes([0x01, 0x21, 0x01, 0x16, 0x21, 0x02, 0x16, 0x00])
    total = 1 + len(body) + 4
    tp = (total + 3) & ~3
    body += b"\0" * (tp - total)
    hdr = bytes([tp // 4 - 1]) + body
    hdr += struct.pack("<I", zlib.crc32(hdr))
    blk = hdr + lvl[2]
    unp = len(blk) + 4
    blk += b"\0" * ((4 - len(blk) % 4) % 4) + struct.pack("<I", zlib.crc32(plain))
    out.append(("xz-chain-second-props-len2", XZ, _xz_wrap(blk, [(unp, len(plain))], c.CHECK_CRC32), {}))
    return out


def all_cases():
    return (valid_lzma2_cases() + valid_lzma_cases() + hand_encoded_cases() + truncation_and_corruption_cases() +
            xz_cases() + xz_chain_cases())
