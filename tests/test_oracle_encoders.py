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


# The one golden COMPRESSED vector the reference holds whose producer is its own encoder: tests/lzma.rs:197-207
# (decompress_empty_world) and src/decode/stream.rs:393,444 use the exact bytes lzma_compress emits for empty input --
# 13-byte header with unknown size, the end marker (is_match=1, is_rep=0, len = 2 via choice=0 + 3-bit tree, pos_slot 63,
# 26 direct bits, 4 align bits, all ones) and the 5-byte flush.  It pins the header writer, the range encoder's carry /
# shift_low / flush logic and every probability model the marker touches.
EMPTY_WORLD = b"\x5d\x00\x00\x80\x00\xff\xff\xff\xff\xff\xff\xff\xff\x00\x83\xff\xfb\xff\xff\xc0\x00\x00\x00"


def test_encoder_known_answer_empty_world():
    assert oracle.lzma_compress(b"") == EMPTY_WORLD
    assert corpus.dumb_lzma(b"") == EMPTY_WORLD
    # and the reference's hello-world vector (written by a real LZMA encoder, not by lzma_compress) still decodes
    hello = (b"\x5d\x00\x00\x80\x00\xff\xff\xff\xff\xff\xff\xff\xff\x00\x24\x19\x49\x98\x6f\x10\x19\xc6\xd7\x31\xeb\x36"
             b"\x50\xb2\x98\x48\xff\xfe\xa5\xb0\x00")
    assert oracle.lzma_decompress(hello).out == b"Hello world\n"
    assert oracle.lzma_compress(b"Hello world\n") != hello  # literals only: a different (longer) encoding of the same text


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
