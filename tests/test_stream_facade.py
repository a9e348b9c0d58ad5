"""CPU tier: decompress::Stream façade (lzma_rs_b200.Stream) against the reference's own stream tests
(src/decode/stream.rs:348-499), with the device replaced by tests/host_emulation (K1's source as 1-lane C++).
The GPU tier runs the same checks on the device (tests/test_gpu_parity.py::test_stream_facade)."""
import io

import pytest

import corpus
import lzma_rs_b200 as L

SMALL = b"Project Gutenberg's Alice's Adventures in Wonderland, by Lewis Carroll\n"  # tests/files/small.txt (71 bytes)
EMPTY_LZMA = (b"\x5d\x00\x00\x80\x00\xff\xff\xff\xff\xff\xff\xff\xff\x00\x83\xff"
              b"\xfb\xff\xff\xc0\x00\x00\x00")  # stream.rs:393-394


def emul_ctx():
    from test_raw_header import _EmulCtx
    return _EmulCtx()


def check_stream_facade(ctx):
    new = lambda opts=None: L.Stream(io.BytesIO(), opts, ctx)  # noqa: E731
    # test_stream_noop / test_stream_zero (stream.rs:352-373)
    s = new()
    assert s.get_output().getvalue() == b""
    assert s.finish().getvalue() == b""
    s = new()
    s.write_all(b"")
    s.write_all(b"")
    assert s.finish().getvalue() == b""
    # test_bad_header (stream.rs:375-388): the write fails, then finish refuses
    s = new()
    with pytest.raises(L.error.LzmaError, match="LZMA header invalid properties: 255 must be < 225"):
        s.write_all(b"\xff" * 32)
    with pytest.raises(L.error.LzmaError, match="can't finish stream because of previous write error"):
        s.finish()
    # test_stream_incomplete (stream.rs:390-432)
    for end in range(1, len(EMPTY_LZMA)):
        s = new()
        s.write_all(EMPTY_LZMA[:end])
        with pytest.raises(L.error.Error) as ei:
            s.finish()
        want = "failed to read header" if end < 18 else "failed to fill whole buffer"
        assert want in str(ei.value), (end, str(ei.value))
    # test_stream_chunked (stream.rs:434-459): every chunk size
    small_c = corpus.dumb_lzma(SMALL)  # crate::lzma_compress, src/encode/dumbencoder.rs
    for data, expected in ((EMPTY_LZMA, b""), (small_c, SMALL)):
        for chunk in range(1, len(data)):
            s = new()
            for o in range(0, len(data), chunk):
                s.write_all(data[o:o + chunk])
            assert s.finish().getvalue() == expected, chunk
    # test_stream_corrupted (stream.rs:461-472): the façade reports data errors in finish()
    s = new()
    s.write_all(b"corrupted bytes here corrupted bytes here")
    with pytest.raises(L.error.LzmaError, match="beyond output size"):
        s.finish()
    # test_allow_incomplete (stream.rs:474-499): half of the compressed bytes decode to exactly 26 bytes
    half = small_c[:len(small_c) // 2]
    s = new()
    s.write_all(half)
    with pytest.raises(L.error.Error):
        s.finish()
    s = L.Stream.new_with_options(L.decompress.Options(allow_incomplete=True), io.BytesIO())
    s._ctx = ctx
    s.write_all(half)
    assert s.finish().getvalue() == SMALL[:26]
    # a match-bearing stream cut at every 997th byte: allow_incomplete returns a prefix of the plaintext and never errors
    plain = corpus.mixed_text(515, 30_000)
    full = corpus.lzma_alone(plain, dict_size=1 << 16)
    prev = -1
    for cut in range(18, len(full), 997):
        s = L.Stream(io.BytesIO(), L.decompress.Options(allow_incomplete=True), ctx)
        s.write_all(full[:cut])
        got = s.finish().getvalue()
        assert plain.startswith(got) and len(got) >= prev
        prev = len(got)
    assert prev > 0


def test_stream_facade_host_logic():
    check_stream_facade(emul_ctx())
