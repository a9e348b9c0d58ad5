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
    """Stands in for lzma_rs_b200.Context in the CPU tier: same call, executed by tests/host_emulation (K1's source
    compiled as 1-lane C++).  Checks the façade's host logic only; the GPU tier runs the same test on the device."""

    def decompress_one(self, fmt, data, options=None):
        import emul_py
        opt = (options or L.decompress.Options())._native()
        r = emul_py.decode_batch(fmt, [bytes(data)], opt)[0]
        cap = 1 << 16
        while int(r.status["code"]) == -1 and cap < (1 << 28):  # LZB_E_CAPACITY: end-marker .lzma, size unknown
            cap *= 4
            r = emul_py.decode_batch(fmt, [bytes(data)], opt, [cap])[0]
        return L.StreamResult(r.data, r.consumed, r.status, r.display)


def test_raw_decoders_host_logic():
    ctx = _EmulCtx()
    data = corpus.mixed_text(4242, 60_000)
    for blob, size in ((corpus.lzma_alone(data, dict_size=1 << 20), None),
                       (corpus.lzma_alone_known_size(data, dict_size=1 << 16), len(data))):
        rd = io.BytesIO(blob + b"TRAILER")
        params = raw.LzmaParams.read_header(rd)
        dec = raw.LzmaDecoder(params, None, ctx)
        out = io.BytesIO()
        if size is None:  # end marker followed by more bytes: lzma.rs:374-381
            with pytest.raises(L.error.LzmaError, match="end-of-stream marker but more bytes"):
                dec.decompress(rd, out)
        else:
            dec.decompress(rd, out)
            # known size: the decoder stops at the last byte it needs; liblzma's end marker stays unread (lzma.rs:442-445)
            assert out.getvalue() == data and rd.read() == (blob + b"TRAILER")[oracle.lzma_decompress(blob + b"TRAILER").consumed:]
        with pytest.raises(L.error.InternalError, match="reset"):
            dec.decompress(blob[13:])
        dec.reset()
        assert dec.decompress(blob[13:]) == data
    blob = corpus.lzma_alone_known_size(data, dict_size=1 << 16)
    want = oracle.lzma_decompress(blob[:5] + blob[13:], unpacked_mode=2, provided=len(data) - 10)
    dec = raw.LzmaDecoder(raw.LzmaParams(raw.LzmaProperties(3, 0, 2), 1 << 16, len(data)), None, ctx)
    dec.reset(len(data) - 10)
    with pytest.raises(L.error.LzmaError) as ei:
        dec.decompress(blob[13:])
    assert str(ei.value) == want.display
    with pytest.raises(L.error.InternalError):
        raw.LzmaDecoder(raw.LzmaParams(raw.LzmaProperties(3, 0, 2), 100), None, ctx)
    d2 = raw.Lzma2Decoder(ctx)
    rd = io.BytesIO(corpus.raw_lzma2(data) + b"xyz")
    assert d2.decompress(rd) == data and rd.read() == b"xyz"
