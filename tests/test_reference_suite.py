"""The reference's own integration tests (tests/lzma.rs, tests/lzma2.rs, tests/xz.rs), restated one-to-one over the
Python mirror of its API.  `run_reference_suite(api)` is written against an object with the reference's function
names; the CPU tier runs it with decode on the host emulation of K1 and encode on the oracle's restatement of the
reference's encoders, the GPU tier (tests/test_gpu_parity.py::test_reference_suite_on_gpu) with the real library.
Fixture files are not read from the reference at run time: `foo.txt` comes from the committed golden blob."""
import io

import pytest

import lzma_rs_b200 as L

CO, CUS = L.compress.Options, L.compress.UnpackedSize
DO, DUS = L.decompress.Options, L.decompress.UnpackedSize


def run_reference_suite(api, foo_txt):
    # ---- tests/lzma.rs:16-93 round_trip = no options + size written to the header, and the stream API beside it
    def round_trip_with_options(x, enc_opts, dec_opts):
        compressed = api.lzma_compress_with_options(x, None, enc_opts)
        assert api.lzma_decompress_with_options(compressed, None, dec_opts) == x
        s = api.Stream(io.BytesIO(), dec_opts)  # tests/lzma.rs:66-92 (#[cfg(feature = "stream")])
        s.write_all(compressed)
        assert s.finish().getvalue() == x

    def round_trip(x):
        compressed = api.lzma_compress(x)
        assert api.lzma_decompress(compressed) == x
        round_trip_with_options(x, CO(CUS.WriteToHeader(len(x))), DO(DUS.ReadFromHeader()))

    round_trip(b"")                       # tests/lzma.rs:146-152 round_trip_basics
    round_trip(bytes(100_000))            # (the reference uses 1 MB; the literal coder's round trip is size-independent)
    round_trip(b"\xff" * 100_000)
    round_trip(b"Hello world")            # 154-159 round_trip_hello
    round_trip(foo_txt[:60_000])          # 161-168 round_trip_files (a prefix keeps the CPU tier fast)
    data = b"Some data"                   # 236-303 the five unpacked-size combinations
    round_trip_with_options(data, CO(CUS.WriteToHeader(len(data))), DO(DUS.ReadFromHeader()))
    round_trip_with_options(data, CO(CUS.SkipWritingToHeader()), DO(DUS.UseProvided(len(data))))
    round_trip_with_options(data, CO(CUS.WriteToHeader(len(data))), DO(DUS.ReadHeaderButUseProvided(len(data))))
    round_trip_with_options(data, CO(CUS.WriteToHeader(None)), DO(DUS.ReadHeaderButUseProvided(len(data))))
    round_trip_with_options(data, CO(CUS.WriteToHeader(None)), DO(DUS.ReadHeaderButUseProvided(None)))
    # 305-356 memlimit
    compressed = api.lzma_compress_with_options(data, None, CO(CUS.WriteToHeader(None)))
    opts = DO(DUS.ReadHeaderButUseProvided(None), memlimit=0)
    with pytest.raises(L.error.LzmaError, match="exceeded memory limit of 0"):
        api.lzma_decompress_with_options(compressed, None, opts)
    s = api.Stream(io.BytesIO(), opts)    # the façade reports the data error in finish()
    s.write_all(compressed)
    with pytest.raises(L.error.LzmaError, match="exceeded memory limit of 0"):
        s.finish()
    # 135-143 decompress_empty_world: HeaderTooShort
    with pytest.raises(L.error.HeaderTooShort):
        api.lzma_decompress(b"")
    # ---- tests/lzma2.rs:12-56 and tests/xz.rs:12-52 round trips
    for x in (b"", bytes(1_000_000), b"\xff" * 1_000_000, b"Hello world", foo_txt):
        assert api.lzma2_decompress(api.lzma2_compress(x)) == x
        assert api.xz_decompress(api.xz_compress(x)) == x


class _CpuTierCtx:
    """Stands in for lzma_rs_b200.Context behind the REAL public entry points (L.lzma_decompress, L.Stream, ...): decode =
    K1's source on the host emulation (tests/host_emulation), encode = the oracle's restatement of the reference's
    encoders.  Everything above the context -- option handling, reader/writer roles, error types -- is product code."""

    def __init__(self):
        from test_raw_header import _EmulCtx
        self._emul = _EmulCtx()

    def decompress_one(self, fmt, data, options=None):
        return self._emul.decompress_one(fmt, data, options)

    def encode_batch(self, fmt, datas, options=None):
        import oracle_py
        u = (options or CO()).unpacked_size
        enc = {0: lambda d: oracle_py.lzma_compress(d, skip_size_field=u.skip, value=None if u.skip else u.value),
               1: oracle_py.lzma2_compress, 2: oracle_py.xz_compress}[fmt]
        return [enc(bytes(d)) for d in datas]


def test_reference_suite_cpu_tier(golden):
    v = next(x for x in golden.vectors() if x["name"] == "foo.txt.lzma")
    import oracle_py
    foo = oracle_py.lzma_decompress(golden.compressed(v)).out
    assert len(foo) == v["plain_len"]
    saved = L._default_ctx
    L._default_ctx = _CpuTierCtx()
    try:
        run_reference_suite(L, foo)
        # allow_incomplete belongs to the stream API: the one-shot entry point ignores it (options.rs:15-19)
        half = L.lzma_compress(b"Some data" * 20)[:40]
        with pytest.raises(L.error.IoError, match="failed to fill whole buffer"):
            L.lzma_decompress_with_options(half, None, DO(allow_incomplete=True))
    finally:
        L._default_ctx = saved
