"""CPU tier: the oracle's restatement of the reference's encoders (oracle/lzma_oracle_enc.c; src/encode/*).  The
reference's tests hold no golden compressed bytes (parity UNPINNED, see that file's header): like tests/lzma.rs:16-28,
tests/lzma2.rs and tests/xz.rs:30-52 these are round trips -- through the decode oracle AND liblzma -- plus agreement
with the independent Python twin used to build test corpora (tests/corpus.py), the exact-size function of the C ABI,
and the one indirect known answer the reference has (src/decode/stream.rs:474-499)."""
import ctypes as C
import lzma

import pytest

import corpus
import oracle_py as oracle
from lzma_rs_b200 import _native

SMALL = b"Project Gutenberg's Alice's Adventures in Wonderland, by Lewis Carroll\n"
INPUTS = [b"", b"a", b"Hello world", SMALL, bytes(1 << 20), b"\xff" * (1 << 20), corpus.mixed_text(31, 200_000),
          bytes(65535), bytes(65536), bytes(65537), corpus.mixed_text(32, 3 * 65536)]


@pytest.mark.parametrize("i", range(len(INPUTS)))
def test_encoders_round_trip(i):
    d = INPUTS[i]
    c = oracle.lzma_compress(d)  # lzma_compress: unknown size + end marker
    assert c == corpus.dumb_lzma(d)
    assert oracle.lzma_decompress(c).out == d and lzma.decompress(c, format=lzma.FORMAT_ALONE) == d
    c = oracle.lzma_compress(d, value=len(d))  # WriteToHeader(Some(n)): no marker
    assert c == corpus.dumb_lzma(d, unpacked_in_header=len(d)) and oracle.lzma_decompress(c).out == d
    c = oracle.lzma_compress(d, skip_size_field=True)  # SkipWritingToHeader: 5-byte header
    assert oracle.lzma_decompress(c, unpacked_mode=2, provided=len(d)).out == d
    l2 = oracle.lzma2_compress(d)
    assert l2 == corpus.stored_lzma2(d) and oracle.lzma2_decompress(l2).out == d
    assert lzma.decompress(l2, format=lzma.FORMAT_RAW, filters=[{"id": lzma.FILTER_LZMA2, "dict_size": 1 << 23}]) == d
    x = oracle.xz_compress(d)
    assert oracle.xz_decompress(x).out == d and lzma.decompress(x, format=lzma.FORMAT_XZ) == d
    lib = _native.load()  # lzb_encode_bound is pure host code: exact for the stored-chunk formats
    assert lib.lzb_encode_bound(1, None, len(d)) == len(l2) and lib.lzb_encode_bound(2, None, len(d)) == len(x)
    assert lib.lzb_encode_bound(0, None, len(d)) >= len(oracle.lzma_compress(d))


def test_stream_known_answer():
    c = oracle.lzma_compress(SMALL)
    got = oracle.lzma_decompress(c[:len(c) // 2])  # truncated: error, but the window held 26 bytes (stream.rs:497-498)
    assert not got.ok
    import io
    import lzma_rs_b200 as L
    from test_raw_header import _EmulCtx
    s = L.Stream(io.BytesIO(), L.decompress.Options(allow_incomplete=True), _EmulCtx())
    s.write_all(c[:len(c) // 2])
    assert s.finish().getvalue() == SMALL[:26]
