"""CPU tier: (1) the oracle's restatement of decompress::Stream (oracle/lzma_oracle.c, lzo_stream_*) pinned on the
reference's own stream tests (src/decode/stream.rs:348-499: known answers incl. allow_incomplete -> small.txt[..26]);
(2) the product's buffering facade (lzma_rs_b200.Stream over the batch path, here on the host emulation of K1) against
that oracle: same final outcome for every chunking, truncation point and option, with the one documented divergence."""
import io
import random

import pytest

import corpus
import lzma_rs_b200 as L
import oracle_py as oracle
from test_stream_facade import EMPTY_LZMA, SMALL, emul_ctx


def test_stream_oracle_on_reference_tests():
    # test_stream_noop / test_stream_zero
    assert oracle.Stream().finish().out == b""
    s = oracle.Stream()
    assert s.write_all(b"") is None and s.write_all(b"") is None and s.finish().out == b""
    # test_bad_header: the write fails with the header error, finish then refuses
    s = oracle.Stream()
    err = s.write_all(b"\xff" * 32)
    assert err is not None and "LZMA header invalid properties: 255 must be < 225" in err.display
    assert "can't finish stream because of previous write error" in s.finish().display
    # test_stream_incomplete: every prefix of the 23-byte empty stream
    for end in range(1, len(EMPTY_LZMA)):
        s = oracle.Stream()
        assert s.write_all(EMPTY_LZMA[:end]) is None
        r = s.finish()
        assert ("failed to read header" if end < 18 else "failed to fill whole buffer") in r.display, (end, r.display)
    # test_stream_chunked: every chunk size
    small_c = oracle.lzma_compress(SMALL)
    for data, expected in ((EMPTY_LZMA, b""), (small_c, SMALL)):
        for chunk in range(1, len(data)):
            s = oracle.Stream()
            for o in range(0, len(data), chunk):
                assert s.write_all(data[o:o + chunk]) is None
            r = s.finish()
            assert r.ok and r.out == expected, (chunk, r.display)
    # test_stream_corrupted: the error comes out of write_all, finish refuses
    s = oracle.Stream()
    err = s.write_all(b"corrupted bytes here corrupted bytes here")
    assert err is not None and "beyond output size" in err.display
    assert "previous write error" in s.finish().display
    # test_allow_incomplete: the known answer
    half = small_c[:len(small_c) // 2]
    s = oracle.Stream()
    assert s.write_all(half) is None and not s.finish().ok
    s = oracle.Stream(allow_incomplete=True)
    assert s.write_all(half) is None
    r = s.finish()
    assert r.ok and r.out == SMALL[:26]
    # tests/lzma.rs:305-356 memlimit through the stream API
    c = oracle.lzma_compress(b"Some data")
    s = oracle.Stream(unpacked_mode=1, memlimit=0)
    err = s.write_all(c)
    assert err is not None and "exceeded memory limit of 0" in err.display
    assert "previous write error" in s.finish().display


def _oracle_outcome(data, chunk, **opts):
    s = oracle.Stream(**opts)
    err = None
    for o in range(0, len(data), chunk):
        err = s.write_all(data[o:o + chunk])
        if err is not None:
            # WriteZero is std's write_all giving up because the finished decoder accepts nothing more (bytes behind
            # the end of a known-size stream, tests/lzma.rs:68-84); the stream itself is fine and finish() succeeds
            if "failed to write whole buffer" in err.display:
                err = None
            break
    r = s.finish()
    if err is not None:  # the reference reports the failure in write; the sink's content is lost with the stream
        return False, err.display
    return r.ok, (r.out if r.ok else r.display)


def _facade_outcome(ctx, data, chunk, unpacked_mode=0, provided=None, memlimit=None, allow_incomplete=False):
    us = L.decompress.UnpackedSize(unpacked_mode, provided)
    s = L.Stream(io.BytesIO(), L.decompress.Options(us, memlimit, allow_incomplete), ctx)
    try:
        for o in range(0, len(data), chunk):
            s.write_all(data[o:o + chunk])
        return True, s.finish().getvalue()
    except L.error.Error as e:
        return False, str(e)


def _same_error(a, b):
    """The reference wraps data errors met inside write into io::Error with Debug formatting, e.g.
    `io error: LzmaError("LZ distance 5 is beyond output size 2")`; the facade raises the error itself."""
    core = lambda t: t.split(": ", 1)[-1].replace('LzmaError("', "").replace('")', "").replace("lzma error: ", "")  # noqa: E731
    return core(a) == core(b) or core(a) in b or core(b) in a


def test_facade_matches_stream_oracle():
    ctx = emul_ctx()
    rnd = random.Random(2026)
    plain = corpus.mixed_text(77, 9000)
    streams = [corpus.lzma_alone(plain, dict_size=4096), corpus.lzma_alone_known_size(plain, dict_size=1 << 16),
               corpus.dumb_lzma(SMALL), corpus.dumb_lzma(SMALL, unpacked_in_header=len(SMALL)), EMPTY_LZMA]
    n = 0
    for data in streams:
        cuts = sorted({len(data)} | {rnd.randrange(1, len(data)) for _ in range(25)} | set(range(1, min(40, len(data)))))
        for cut in cuts:
            for allow in (False, True):
                chunk = rnd.choice([1, 2, 7, 19, 20, 21, 64, 1000, 1 << 20])
                want = _oracle_outcome(data[:cut], chunk, allow_incomplete=allow)
                got = _facade_outcome(ctx, data[:cut], chunk, allow_incomplete=allow)
                n += 1
                assert want[0] == got[0], (len(data), cut, chunk, allow, want[:1], got)
                if want[0]:
                    assert want[1] == got[1], (len(data), cut, chunk, allow)
                else:
                    assert _same_error(want[1], got[1]), (cut, chunk, allow, want[1], got[1])
    assert n > 300


def test_facade_matches_stream_oracle_on_generated_streams():
    """Valid .lzma streams of every shape the symbol-level generator produces (all property combinations incl. lc+lp > 4,
    known and unknown sizes, with and without end marker), complete or cut at a random byte, any write chunking."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_soak
    ctx = emul_ctx()
    rnd = random.Random(14)
    n = 0
    for _ in range(700):
        fmt, data = fuzz_soak.lzma_structured(rnd, corpus)
        if fmt != 0 or len(data) < 2 or not oracle.lzma_decompress(data).ok:
            continue
        for rep in range(2):
            cut = len(data) if rep == 0 else rnd.randrange(1, len(data) + 1)
            allow = rnd.random() < 0.6
            chunk = rnd.choice([1, 3, 19, 20, 21, 100, 1 << 20])
            want = _oracle_outcome(data[:cut], chunk, allow_incomplete=allow)
            got = _facade_outcome(ctx, data[:cut], chunk, allow_incomplete=allow)
            n += 1
            assert want[0] == got[0], (data.hex(), cut, chunk, allow, want, got)
            assert want[1] == got[1] if want[0] else _same_error(want[1], got[1]), (data.hex(), cut, chunk, allow)
    assert n > 300
