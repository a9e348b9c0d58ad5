"""Seeded synthetic corpora and hand encoders for tests and bench.py (test/bench infrastructure).

Valid match-bearing streams come from liblzma (Python `lzma`), the same role `rust-lzma`/`xz2` play in the
reference's tests (tests/lzma.rs:1,111-114; fuzz/fuzz_targets/interop_xz_decode.rs).  Adversarial and malformed
streams come from `LzmaEncoder`, a small range *encoder* whose arithmetic mirrors the decoder contract
(reference: src/encode/rangecoder.rs:36-106 for the carry/cache scheme, src/decode/lzma.rs:278-393 for the symbol
grammar).  Container writers (LZMA2 chunks, .xz) follow src/encode/{lzma2,xz}.rs and the .xz format.
"""
import lzma
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

# ------------------------------------------------------------------------------------------------
# data model: "mixed literal/match" text
# ------------------------------------------------------------------------------------------------
_VOCAB = None


def _vocab():
    global _VOCAB
    if _VOCAB is None:
        rng = np.random.default_rng(12345)
        lens = rng.integers(3, 11, size=512)  # 2..9 letters + trailing space
        table = np.zeros((512, 10), dtype=np.uint8)
        for i, ln in enumerate(lens):
            table[i, : ln - 1] = rng.integers(97, 123, size=ln - 1)
            table[i, ln - 1] = 32
        _VOCAB = (table, lens.astype(np.int64))
    return _VOCAB


def mixed_text(seed, n):
    """n bytes: ~70 % segments of vocabulary words, ~30 % segments of random bytes (vectorised, seeded)."""
    if n == 0:
        return b""
    table, wlen = _vocab()
    rng = np.random.default_rng(seed)
    out = np.empty(0, dtype=np.uint8)
    pieces = []
    total = 0
    while total < n:
        nseg = max(16, (n - total) // 80 + 8)
        is_word = rng.random(nseg) < 0.7
        nwords = rng.integers(4, 40, size=nseg)
        rlen = rng.integers(8, 200, size=nseg)
        # tokens: each word of a word-segment, or one whole random run
        tok_per_seg = np.where(is_word, nwords, 1)
        seg_of_tok = np.repeat(np.arange(nseg), tok_per_seg)
        tok_is_word = is_word[seg_of_tok]
        ntok = len(seg_of_tok)
        widx = rng.integers(0, 512, size=ntok)
        tok_len = np.where(tok_is_word, wlen[widx], rlen[seg_of_tok])
        tot = int(tok_len.sum())
        start = np.cumsum(tok_len) - tok_len
        tok_of_byte = np.repeat(np.arange(ntok), tok_len)
        j = np.arange(tot) - start[tok_of_byte]
        rnd = rng.integers(0, 256, size=tot, dtype=np.uint8)
        wb = table[widx[tok_of_byte], np.minimum(j, 9)]
        buf = np.where(tok_is_word[tok_of_byte], wb, rnd).astype(np.uint8)
        pieces.append(buf)
        total += tot
    out = np.concatenate(pieces)[:n]
    return out.tobytes()


# ------------------------------------------------------------------------------------------------
# liblzma-backed encoders (valid streams)
# ------------------------------------------------------------------------------------------------
def raw_lzma2(data, dict_size=1 << 18, lc=3, lp=0, pb=2, preset=6):
    """Raw LZMA2 stream: exactly the bytes lzma_rs::lzma2_decompress consumes (chunks + 0x00)."""
    f = [{"id": lzma.FILTER_LZMA2, "preset": preset, "dict_size": dict_size, "lc": lc, "lp": lp, "pb": pb}]
    return lzma.compress(data, format=lzma.FORMAT_RAW, filters=f)


def lzma_alone(data, dict_size=1 << 16, lc=3, lp=0, pb=2, preset=6):
    """.lzma ("alone") file; liblzma writes unknown size (-1) + end marker."""
    f = [{"id": lzma.FILTER_LZMA1, "preset": preset, "dict_size": dict_size, "lc": lc, "lp": lp, "pb": pb}]
    return lzma.compress(data, format=lzma.FORMAT_ALONE, filters=f)


def lzma_alone_known_size(data, **kw):
    """.lzma with the real size in the header and NO end marker requirement (marker still present but unread,
    reference leniency (e): lzma.rs:442-445 stops at the size)."""
    c = bytearray(lzma_alone(data, **kw))
    c[5:13] = struct.pack("<Q", len(data))
    return bytes(c)


# --- CRC-64/XZ (test-side, independent of the oracle and the product) ---
_CRC64_TAB = None


def crc64_xz(data):
    global _CRC64_TAB
    if _CRC64_TAB is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0xC96C5795D7870F42 if c & 1 else c >> 1
            t.append(c)
        _CRC64_TAB = t
    c = 0xFFFFFFFFFFFFFFFF
    for b in data:
        c = _CRC64_TAB[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFFFFFFFFFF


def _multibyte(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v == 0:
            out.append(b)
            return bytes(out)
        out.append(0x80 | b)


CHECK_NONE, CHECK_CRC32, CHECK_CRC64, CHECK_SHA256 = 0x00, 0x01, 0x04, 0x0A


def xz_block(payload_lzma2, plain, check=CHECK_CRC32, with_sizes=False, dict_prop=0x16, nfilters=1):
    """One .xz block (header + LZMA2 payload + padding + check).  Returns (bytes, unpadded_size)."""
    flags = (nfilters - 1) | (0xC0 if with_sizes else 0)
    body = bytes([flags])
    if with_sizes:
        body += _multibyte(len(payload_lzma2)) + _multibyte(len(plain))
    for _ in range(nfilters):
        body += bytes([0x21, 0x01, dict_prop])
    total = 1 + len(body) + 4
    total_padded = (total + 3) & ~3
    body += b"\0" * (total_padded - total)
    hdr = bytes([total_padded // 4 - 1]) + body
    hdr += struct.pack("<I", zlib.crc32(hdr))
    blk = hdr + payload_lzma2
    unpadded = len(blk)
    blk += b"\0" * ((4 - len(blk) % 4) % 4)
    if check == CHECK_CRC32:
        blk += struct.pack("<I", zlib.crc32(plain))
        unpadded += 4
    elif check == CHECK_CRC64:
        blk += struct.pack("<Q", crc64_xz(plain))
        unpadded += 8
    elif check == CHECK_SHA256:
        import hashlib
        blk += hashlib.sha256(plain).digest()
        unpadded += 32
    return blk, unpadded


def xz_file(data, block_size=1 << 18, check=CHECK_CRC32, with_sizes=False, dict_size=1 << 20, preset=6):
    """Multi-block .xz (one LZMA2 stream per block), like `xz --block-size=N --check=...`."""
    flags = bytes([0, check])
    out = bytearray(b"\xfd7zXZ\0" + flags + struct.pack("<I", zlib.crc32(flags)))
    records = []
    for off in range(0, len(data), block_size) if data else []:
        plain = data[off:off + block_size]
        payload = raw_lzma2(plain, dict_size=dict_size, preset=preset)
        blk, unpadded = xz_block(payload, plain, check, with_sizes)
        out += blk
        records.append((unpadded, len(plain)))
    idx = bytearray(b"\0" + _multibyte(len(records)))
    for u, p in records:
        idx += _multibyte(u) + _multibyte(p)
    idx += b"\0" * ((4 - len(idx) % 4) % 4)
    idx += struct.pack("<I", zlib.crc32(bytes(idx)))
    out += idx
    footer_body = struct.pack("<I", len(idx) // 4 - 1) + flags
    out += struct.pack("<I", zlib.crc32(footer_body)) + footer_body + b"YZ"
    return bytes(out)


# ------------------------------------------------------------------------------------------------
# hand range encoder (adversarial / malformed / reference-encoder-shaped streams)
# ------------------------------------------------------------------------------------------------
class RangeEncoder:
    """Carry-propagating LZMA range encoder (same byte accounting as src/encode/rangecoder.rs:20-106)."""

    def __init__(self):
        self.low = 0
        self.range = 0xFFFFFFFF
        self.cache = 0
        self.cachesz = 1
        self.out = bytearray()

    def _shift_low(self):
        if self.low < 0xFF000000 or self.low > 0xFFFFFFFF:
            carry = self.low >> 32
            tmp = self.cache
            while True:
                self.out.append((tmp + carry) & 0xFF)
                tmp = 0xFF
                self.cachesz -= 1
                if self.cachesz == 0:
                    break
            self.cache = (self.low >> 24) & 0xFF
        self.cachesz += 1
        self.low = (self.low << 8) & 0xFFFFFFFF

    def _normalize(self):
        while self.range < 0x01000000:
            self.range = (self.range << 8) & 0xFFFFFFFF
            self._shift_low()

    def bit(self, probs, idx, bit):
        p = probs[idx]
        bound = (self.range >> 11) * p
        if bit:
            probs[idx] = p - (p >> 5)
            self.low += bound
            self.range -= bound
        else:
            probs[idx] = p + ((0x800 - p) >> 5)
            self.range = bound
        self._normalize()

    def direct(self, value, nbits):
        for i in range(nbits - 1, -1, -1):
            self.range >>= 1
            if (value >> i) & 1:
                self.low += self.range
            self._normalize()

    def finish(self):
        for _ in range(5):
            self._shift_low()
        return bytes(self.out)


class LzmaEncoder:
    """Symbol-level LZMA encoder: the caller dictates the parse (literal / match / rep / shortrep / end marker),
    the encoder keeps probabilities, `state` and `rep[]` in step with the decoder contract."""

    def __init__(self, lc=3, lp=0, pb=2, history=b""):
        self.lc, self.lp, self.pb = lc, lp, pb
        self.rc = RangeEncoder()
        self.hist = bytearray(history)  # bytes since the last dict reset (the decoder's window)
        self.reset_state()

    def reset_state(self, lc=None, lp=None, pb=None):
        if lc is not None:
            self.lc, self.lp, self.pb = lc, lp, pb
        n = 0x300 << (self.lc + self.lp)
        self.lit = [0x400] * n
        self.is_match = [0x400] * 192
        self.is_rep = [0x400] * 12
        self.is_rep_g0 = [0x400] * 12
        self.is_rep_g1 = [0x400] * 12
        self.is_rep_g2 = [0x400] * 12
        self.is_rep_0long = [0x400] * 192
        self.pos_slot = [[0x400] * 64 for _ in range(4)]
        self.pos_dec = [0x400] * 115
        self.align = [0x400] * 16
        self.len_dec = self._new_len()
        self.rep_len_dec = self._new_len()
        self.state = 0
        self.rep = [0, 0, 0, 0]

    def new_chunk(self):
        """LZMA2: a fresh range coder per chunk (reference lzma2.rs:190)."""
        self.rc = RangeEncoder()

    @staticmethod
    def _new_len():
        return {"choice": [0x400, 0x400], "low": [[0x400] * 8 for _ in range(16)],
                "mid": [[0x400] * 8 for _ in range(16)], "high": [0x400] * 256}

    def _pos_state(self):
        return len(self.hist) & ((1 << self.pb) - 1)

    def _tree(self, probs, nbits, value):
        m = 1
        for i in range(nbits - 1, -1, -1):
            b = (value >> i) & 1
            self.rc.bit(probs, m, b)
            m = (m << 1) | b

    def _rtree(self, probs, offset, nbits, value):
        m = 1
        for i in range(nbits):
            b = (value >> i) & 1
            self.rc.bit(probs, offset + m, b)
            m = (m << 1) | b

    def _len(self, dec, length):  # length = real length - 2
        ps = self._pos_state()
        if length < 8:
            self.rc.bit(dec["choice"], 0, 0)
            self._tree(dec["low"][ps], 3, length)
        elif length < 16:
            self.rc.bit(dec["choice"], 0, 1)
            self.rc.bit(dec["choice"], 1, 0)
            self._tree(dec["mid"][ps], 3, length - 8)
        else:
            self.rc.bit(dec["choice"], 0, 1)
            self.rc.bit(dec["choice"], 1, 1)
            self._tree(dec["high"], 8, length - 16)

    def literal(self, byte):
        ps = self._pos_state()
        self.rc.bit(self.is_match, (self.state << 4) + ps, 0)
        prev = self.hist[-1] if self.hist else 0
        row = ((len(self.hist) & ((1 << self.lp) - 1)) << self.lc) + (prev >> (8 - self.lc))
        base = row * 0x300
        m = 1
        i = 7
        if self.state >= 7:
            mb = self.hist[-(self.rep[0] + 1)]
            while i >= 0:
                match_bit = (mb >> i) & 1
                b = (byte >> i) & 1
                self.rc.bit(self.lit, base + ((1 + match_bit) << 8) + m, b)
                m = (m << 1) | b
                i -= 1
                if match_bit != b:
                    break
        while i >= 0:
            b = (byte >> i) & 1
            self.rc.bit(self.lit, base + m, b)
            m = (m << 1) | b
            i -= 1
        self.hist.append(byte)
        s = self.state
        self.state = 0 if s < 4 else (s - 3 if s < 10 else s - 6)

    def _copy(self, length, dist):
        for _ in range(length):
            self.hist.append(self.hist[-dist])

    def _distance(self, dist0, len_minus2):
        """dist0 = distance - 1 (the value stored in rep[0])."""
        ls = min(len_minus2, 3)
        if dist0 < 4:
            slot = dist0
        else:
            n = dist0.bit_length() - 1
            slot = (n << 1) | ((dist0 >> (n - 1)) & 1)
        self._tree(self.pos_slot[ls], 6, slot)
        if slot >= 4:
            nd = (slot >> 1) - 1
            base = (2 | (slot & 1)) << nd
            rem = dist0 - base
            if slot < 14:
                self._rtree(self.pos_dec, base - slot, nd, rem)
            else:
                self.rc.direct(rem >> 4, nd - 4)
                self._rtree(self.align, 0, 4, rem & 15)

    def match(self, length, dist, check=True):
        """New match: length 2..273, dist >= 1 (dist-1 goes to rep[0]).  check=False allows invalid distances."""
        ps = self._pos_state()
        self.rc.bit(self.is_match, (self.state << 4) + ps, 1)
        self.rc.bit(self.is_rep, self.state, 0)
        self.rep = [dist - 1, self.rep[0], self.rep[1], self.rep[2]]
        self._len(self.len_dec, length - 2)
        self.state = 7 if self.state < 7 else 10
        self._distance(dist - 1, length - 2)
        if check:
            self._copy(length, dist)

    def end_marker(self):
        ps = self._pos_state()
        self.rc.bit(self.is_match, (self.state << 4) + ps, 1)
        self.rc.bit(self.is_rep, self.state, 0)
        self._len(self.len_dec, 0)
        self.state = 7 if self.state < 7 else 10
        self._distance(0xFFFFFFFF, 0)

    def shortrep(self):
        ps = self._pos_state()
        self.rc.bit(self.is_match, (self.state << 4) + ps, 1)
        self.rc.bit(self.is_rep, self.state, 1)
        self.rc.bit(self.is_rep_g0, self.state, 0)
        self.rc.bit(self.is_rep_0long, (self.state << 4) + ps, 0)
        self.state = 9 if self.state < 7 else 11
        self._copy(1, self.rep[0] + 1)

    def rep_match(self, idx, length, check=True):
        """Rep match with rep[idx], length 2..273."""
        ps = self._pos_state()
        self.rc.bit(self.is_match, (self.state << 4) + ps, 1)
        self.rc.bit(self.is_rep, self.state, 1)
        if idx == 0:
            self.rc.bit(self.is_rep_g0, self.state, 0)
            self.rc.bit(self.is_rep_0long, (self.state << 4) + ps, 1)
        else:
            self.rc.bit(self.is_rep_g0, self.state, 1)
            if idx == 1:
                self.rc.bit(self.is_rep_g1, self.state, 0)
            else:
                self.rc.bit(self.is_rep_g1, self.state, 1)
                self.rc.bit(self.is_rep_g2, self.state, idx - 2)
            d = self.rep[idx]
            for i in range(idx, 0, -1):
                self.rep[i] = self.rep[i - 1]
            self.rep[0] = d
        self._len(self.rep_len_dec, length - 2)
        self.state = 8 if self.state < 7 else 11
        if check:
            self._copy(length, self.rep[0] + 1)

    def finish(self):
        return self.rc.finish()


def props_byte(lc, lp, pb):
    return (pb * 5 + lp) * 9 + lc


def lzma_header(lc=3, lp=0, pb=2, dict_size=0x800000, unpacked=None):
    """13-byte .lzma header (src/decode/lzma.rs:96-161); unpacked=None -> 0xFFFF_FFFF_FFFF_FFFF."""
    return bytes([props_byte(lc, lp, pb)]) + struct.pack("<I", dict_size) + struct.pack(
        "<Q", 0xFFFFFFFFFFFFFFFF if unpacked is None else unpacked)


def dumb_lzma(data, unpacked_in_header=None, write_size_field=True, end_marker=None):
    """Literal-only .lzma like the reference's own encoder (src/encode/dumbencoder.rs:24-123): lc3 lp0 pb2,
    dict 0x800000; end marker iff the header says "unknown size"."""
    enc = LzmaEncoder(3, 0, 2)
    for b in data:
        enc.literal(b)
    if end_marker is None:
        end_marker = write_size_field and unpacked_in_header is None
    if end_marker:
        enc.end_marker()
    hdr = bytes([props_byte(3, 0, 2)]) + struct.pack("<I", 0x800000)
    if write_size_field:
        hdr += struct.pack("<Q", 0xFFFFFFFFFFFFFFFF if unpacked_in_header is None else unpacked_in_header)
    return hdr + enc.finish()


def stored_lzma2(data):
    """Stored-chunk-only LZMA2 like the reference's own encoder (src/encode/lzma2.rs:4-26)."""
    out = bytearray()
    for off in range(0, len(data), 0x10000):
        piece = data[off:off + 0x10000]
        out += bytes([1]) + struct.pack(">H", len(piece) - 1) + piece
    out.append(0)
    return bytes(out)


def lzma2_chunk(payload, unpacked_len, control, props=None):
    """One compressed LZMA2 chunk header + payload.  control in {0x80,0xA0,0xC0,0xE0}."""
    u = unpacked_len - 1
    hdr = bytes([control | (u >> 16)]) + struct.pack(">H", u & 0xFFFF) + struct.pack(">H", len(payload) - 1)
    if control >= 0xC0:
        hdr += bytes([props])
    return hdr + payload


def rep0_stress_lzma2(total, byte=0x41, lc=3, lp=0, pb=2, chunk_unpacked=1 << 21):
    """BASELINE config 5: `literal; match(dist=1,len=273); rep0-long(len=273)...; tail` as LZMA2 chunks."""
    out = bytearray()
    produced = 0
    first = True
    have_match = False
    enc = LzmaEncoder(lc, lp, pb)
    while produced < total:
        n = min(chunk_unpacked, total - produced)
        enc.new_chunk()
        start = len(enc.hist)
        left = n
        if first:
            enc.literal(byte)
            left -= 1
        while left > 0:
            ln = min(273, left)
            if not have_match:
                if ln >= 2:
                    enc.match(ln, 1)
                    have_match = True
                else:
                    enc.literal(byte)
            elif ln == 1:
                enc.shortrep()
            else:
                enc.rep_match(0, ln)
            left -= ln
        payload = enc.finish()
        assert len(enc.hist) - start == n
        out += lzma2_chunk(payload, n, 0xE0 if first else 0x80, props_byte(lc, lp, pb))
        produced += n
        first = False
    out.append(0)
    return bytes(out)


# ------------------------------------------------------------------------------------------------
# batch corpora for BASELINE configs (bench.py and the full-size GPU tests)
# ------------------------------------------------------------------------------------------------
def build_lzma2_corpus(config_index, n_streams, size_fn, dict_size, distinct=None, threads=None, preset=6):
    """Returns (list of compressed streams, list of plaintexts).  `distinct` < n_streams tiles the first
    `distinct` streams (stated in the bench's config) to bound generation time."""
    distinct = n_streams if distinct is None else min(distinct, n_streams)

    def one(i):
        seed = config_index * 1_000_003 + i
        plain = mixed_text(seed, size_fn(i))
        return raw_lzma2(plain, dict_size=dict_size, preset=preset), plain

    with ThreadPoolExecutor(max_workers=threads) as ex:
        base = list(ex.map(one, range(distinct)))
    comp = [base[i % distinct][0] for i in range(n_streams)]
    plain = [base[i % distinct][1] for i in range(n_streams)]
    return comp, plain


def pack_blob(streams, align=16):
    """Concatenate streams into one uint8 blob with `align`-byte aligned starts; returns (blob, off[n+1], len[n]).
    off[i] is the start of stream i, len[i] its length (off[i+1]-off[i] includes alignment padding)."""
    n = len(streams)
    lens = np.fromiter((len(s) for s in streams), dtype=np.uint64, count=n)
    padded = (lens + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(padded, out=off[1:])
    blob = np.zeros(int(off[-1]) + 64, dtype=np.uint8)
    for i, s in enumerate(streams):
        o = int(off[i])
        blob[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
    return blob, off, lens
