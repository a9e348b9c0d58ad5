"""CPU tier: host-side pieces of the decompress::raw mirror that need no device -- LzmaParams::read_header
(src/decode/lzma.rs:96-161) against the oracle's view of the same headers, and property validation."""
import io

import pytest

import corpus
import oracle_py as oracle
import lzma_rs_b200 as L

raw = L.decompress.raw
US = L.decompress.UnpackedSize


def test_read_header_modes():
    data = b"Hello world\n" * 50
    blob = corpus.lzma_alone_known_size(data, dict_size=1 << 16)
    p = raw.LzmaParams.read_header(io.BytesIO(blob))
    assert (p.properties.lc, p.properties.lp, p.properties.pb, p.dict_size, p.unpacked_size) == (3, 0, 2, 1 << 16, len(data))
    rd = io.BytesIO(blob)
    p = raw.LzmaParams.read_header(rd, L.decompress.Options(US.ReadHeaderButUseProvided(7)))
    assert p.unpacked_size == 7 and rd.tell() == 13
    rd = io.BytesIO(blob)
    p = raw.LzmaParams.read_header(rd, L.decompress.Options(US.UseProvided(None)))
    assert p.unpacked_size is None and rd.tell() == 5
    marker = bytearray(corpus.lzma_alone(data, dict_size=4096))
    marker[1:5] = (100).to_bytes(4, "little")  # header says 100 -> clamped to 0x1000 (lzma.rs:122-126)
    p = raw.LzmaParams.read_header(io.BytesIO(marker))
    assert p.unpacked_size is None and p.dict_size == 0x1000


@pytest.mark.parametrize("n", [0, 1, 4, 5, 12])
def test_read_header_too_short_matches_oracle(n):
    blob = corpus.lzma_alone_known_size(b"abc" * 100)[:n]
    want = oracle.lzma_decompress(blob)
    assert want.kind == 2
    with pytest.raises(L.error.HeaderTooShort) as ei:
        raw.LzmaParams.read_header(io.BytesIO(blob))
    assert str(ei.value) == want.display


def test_read_header_invalid_props_matches_oracle():
    want = oracle.lzma_decompress(b"\xff" * 32)
    with pytest.raises(L.error.LzmaError) as ei:
        raw.LzmaParams.read_header(io.BytesIO(b"\xff" * 32))
    assert str(ei.value) == want.display == "lzma error: LZMA header invalid properties: 255 must be < 225"


def test_properties_validate():
    raw.LzmaProperties(8, 4, 4).validate()
    for bad in ((9, 0, 0), (0, 5, 0), (0, 0, 5)):
        with pytest.raises(AssertionError):
            raw.LzmaProperties(*bad).validate()


class _EmulCtx:
    """Stands in for lzma_rs_b200.Context in the CPU tier: same calls, executed by tests/host_emulation (K1's source
    compiled as 1-lane C++).  Checks the host logic and K1's decode logic; the GPU tier runs the same tests on the device."""

    def decompress_one(self, fmt, data, options=None):
        import emul_py
        opt = (options or L.decompress.Options())._native()
        r = emul_py.decode_batch(fmt, [bytes(data)], opt)[0]
        cap = 1 << 16
        while int(r.status["code"]) == -1 and cap < (1 << 28):  # LZB_E_CAPACITY: end-marker .lzma, size unknown
            cap *= 4
            r = emul_py.decode_batch(fmt, [bytes(data)], opt, [cap])[0]
        return L.StreamResult(r.data, r.consumed, r.status, r.display)

    def raw_new(self, fmt, lc, lp, pb, dict_size):
        import emul_py
        h = emul_py.RawHandle(fmt, lc, lp, pb, dict_size)
        dec = h.decompress
        h.decompress = lambda data, options=None: (lambda r: L.StreamResult(r.data, r.consumed, r.status, r.display))(dec(data, options))
        return h


def raw_decoder_checks(ctx):
    """decompress::raw::{LzmaParams, LzmaDecoder, Lzma2Decoder} against the oracle's decoder OBJECTS, including what the
    reference does when decompress() is called again without reset(): the DecoderState carries over (lzma.rs:597-648,
    lzma2.rs:11-82).  Shared by the CPU tier (host emulation) and the GPU tier."""
    data = corpus.mixed_text(4242, 60_000)
    for blob, size in ((corpus.lzma_alone(data, dict_size=1 << 20), None),
                       (corpus.lzma_alone_known_size(data, dict_size=1 << 16), len(data))):
        rd = io.BytesIO(blob + b"TRAILER")
        params = raw.LzmaParams.read_header(rd)
        assert params.unpacked_size == size and (params.properties.lc, params.properties.lp, params.properties.pb) == (3, 0, 2)
        dec = raw.LzmaDecoder(params, None, ctx)
        ora = oracle.RawDecoder(0, 3, 0, 2, params.dict_size, size)
        out = io.BytesIO()
        want = ora.decompress(blob[13:] + b"TRAILER")
        if size is None:  # end marker followed by more bytes: lzma.rs:374-381
            assert not want.ok
            with pytest.raises(L.error.LzmaError, match="end-of-stream marker but more bytes") as ei:
                dec.decompress(rd, out)
            assert str(ei.value) == want.display
            dec.reset()
            ora.reset()
            assert dec.decompress(blob[13:]) == data == ora.decompress(blob[13:]).out
        else:
            dec.decompress(rd, out)
            # known size: the decoder stops at the last byte it needs; liblzma's end marker stays unread (lzma.rs:442-445)
            assert out.getvalue() == data == want.out and rd.read() == (blob[13:] + b"TRAILER")[want.consumed:]
        # a second decompress() WITHOUT reset(): same bytes, carried probabilities -> whatever the reference makes of it
        for again in range(2):
            want2 = ora.decompress(blob[13:])
            try:
                got2, disp2 = dec.decompress(blob[13:]), ""
            except L.error.Error as e:
                got2, disp2 = None, str(e)
            assert disp2 == want2.display, (size, again, disp2, want2.display)
            if want2.ok:
                assert got2 == want2.out
            else:
                break
        dec.reset()
        ora.reset()
        assert dec.decompress(blob[13:]) == data == ora.decompress(blob[13:]).out
    # two DIFFERENT streams through one decoder without reset: the second one starts from the first one's probabilities,
    # state and rep distances, so it decodes correctly only if it was ENCODED from that state -- two pieces of one encoder
    # run, each flushed with its own range coder, both n bytes long (the expected size is part of the carried state)
    n = 48
    enc = corpus.LzmaEncoder(3, 0, 2)
    for b in b"carry me over, carry me over, ca":
        enc.literal(b)
    enc.match(11, 15)
    for b in b"rry! ":
        enc.literal(b)
    assert len(enc.hist) == n
    part1 = enc.finish()
    enc.rc, enc.hist = corpus.RangeEncoder(), bytearray()  # same models / state / reps, new range coder, empty window
    second = b"over the state, over and over and over again ..."
    assert len(second) == n
    for b in second[:30]:
        enc.literal(b)
    enc.match(6, 11)
    while len(enc.hist) < n:
        enc.literal(second[len(enc.hist)])
    plain2 = bytes(enc.hist)
    part2 = enc.finish()
    dec = raw.LzmaDecoder(raw.LzmaParams(raw.LzmaProperties(3, 0, 2), 1 << 12, n), None, ctx)
    ora = oracle.RawDecoder(0, 3, 0, 2, 1 << 12, n)
    w1, w2 = ora.decompress(part1), ora.decompress(part2)
    assert w1.ok and w2.ok and w2.out == plain2 and len(w1.out) == n
    assert dec.decompress(part1) == w1.out and dec.decompress(part2) == w2.out
    fresh = oracle.RawDecoder(0, 3, 0, 2, 1 << 12, n).decompress(part2)
    assert fresh.out != w2.out or not fresh.ok  # without the carried state the same bytes decode to something else
    # a carried state >= 7 (the previous stream ended in a match) makes the next stream's first literal a MATCHED literal,
    # whose match byte lies in a window that no longer exists: the reference fails with its distance check
    enc = corpus.LzmaEncoder(3, 0, 2)
    for b in b"ends in a match: abcabc":
        enc.literal(b)
    enc.match(4, 3)
    tail_match = enc.finish()
    m = len(enc.hist)
    dec = raw.LzmaDecoder(raw.LzmaParams(raw.LzmaProperties(3, 0, 2), 1 << 12, m), None, ctx)
    ora = oracle.RawDecoder(0, 3, 0, 2, 1 << 12, m)
    assert dec.decompress(tail_match) == ora.decompress(tail_match).out
    want = ora.decompress(part1)
    assert not want.ok and "beyond output size" in want.display
    with pytest.raises(L.error.LzmaError) as ei:
        dec.decompress(part1)
    assert str(ei.value) == want.display
    # same payload against the oracle with an explicit size (UseProvided) and a wrong size
    blob = corpus.lzma_alone_known_size(data, dict_size=1 << 16)
    for short in range(1, 3000):  # find a size the last match overshoots (lzma.rs:513-521); others just stop early
        p = raw.LzmaParams(raw.LzmaProperties(3, 0, 2), 1 << 16, len(data) - short)
        want = oracle.lzma_decompress(blob[:5] + blob[13:], unpacked_mode=2, provided=len(data) - short)
        if want.ok:
            assert raw.LzmaDecoder(p, None, ctx).decompress(blob[13:]) == want.out == data[:len(data) - short]
        else:
            with pytest.raises(L.error.LzmaError) as ei:
                raw.LzmaDecoder(p, None, ctx).decompress(blob[13:])
            assert str(ei.value) == want.display and "Expected unpacked size" in want.display
            break
    else:
        raise AssertionError("no overshooting size found")
    with pytest.raises(AssertionError):
        raw.LzmaDecoder(raw.LzmaParams(raw.LzmaProperties(9, 0, 0), 1 << 16), None, ctx)
    # LzmaParams::new does not clamp the dictionary size (only read_header does): a 100-byte window is a 100-byte window
    small = corpus.lzma_alone_known_size(data[:5000], dict_size=4096)
    p = raw.LzmaParams(raw.LzmaProperties(3, 0, 2), 100, 5000)
    want = oracle.RawDecoder(0, 3, 0, 2, 100, 5000).decompress(small[13:])
    try:
        got, disp = raw.LzmaDecoder(p, None, ctx).decompress(small[13:]), ""
    except L.error.Error as e:
        got, disp = None, str(e)
    assert disp == want.display and (got is None or got == want.out)
    # Lzma2Decoder: trailing bytes stay unread; state (incl. the properties of the last props reset) carries over, so a
    # stream whose first chunk resets nothing (control 0x80) decodes with the previous stream's models
    d2, o2 = raw.Lzma2Decoder(ctx), oracle.RawDecoder(1)
    rd = io.BytesIO(corpus.raw_lzma2(data) + b"xyz")
    assert d2.decompress(rd) == data == o2.decompress(corpus.raw_lzma2(data)).out and rd.read() == b"xyz"
    e = corpus.LzmaEncoder(3, 0, 2)
    for b in b"a chunk that resets nothing":
        e.literal(b)
    tail = corpus.lzma2_chunk(e.finish(), len(e.hist), 0x80) + b"\0"
    want = o2.decompress(tail)
    try:
        got, disp = d2.decompress(tail), ""
    except L.error.Error as ex:
        got, disp = None, str(ex)
    assert disp == want.display and (got is None or got == want.out)
    fresh = oracle.lzma2_decompress(tail)  # a fresh decoder sees lc = lp = pb = 0 tables instead: a different outcome
    assert (fresh.display, fresh.out) != (want.display, want.out)
    d2.reset()
    o2.reset()
    assert d2.decompress(corpus.raw_lzma2(data[:1000])) == data[:1000] == o2.decompress(corpus.raw_lzma2(data[:1000])).out
    want = o2.decompress(tail)
    try:
        got, disp = d2.decompress(tail), ""
    except L.error.Error as ex:
        got, disp = None, str(ex)
    assert disp == want.display and (got is None or got == want.out)


def test_raw_decoders_host_logic():
    raw_decoder_checks(_EmulCtx())
