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


class _CpuTierApi:
    """The reference's entry points with the device replaced: decode = K1's source on the host emulation
    (tests/host_emulation), encode = the oracle's restatement of the reference's encoders."""

    def __init__(self):
        from test_raw_header import _EmulCtx
        self.ctx = _EmulCtx()

    def _dec(self, fmt, data, opts):
        r = self.ctx.decompress_one(fmt, data, opts)
        r.raise_for_status()
        return r.data

    def lzma_decompress(self, d):
        return self._dec(0, d, None)

    def lzma_decompress_with_options(self, d, out, opts):
        return self._dec(0, d, opts)

    def lzma2_decompress(self, d):
        return self._dec(1, d, None)

    def xz_decompress(self, d):
        return self._dec(2, d, None)

    def Stream(self, out, opts=None):
        return L.Stream(out, opts, self.ctx)

    def lzma_compress(self, d):
        import oracle_py
        return oracle_py.lzma_compress(d)

    def lzma_compress_with_options(self, d, out, opts):
        import oracle_py
        u = opts.unpacked_size
        return oracle_py.lzma_compress(d, skip_size_field=u.skip, value=None if u.skip else u.value)

    def lzma2_compress(self, d):
        import oracle_py
        return oracle_py.lzma2_compress(d)

    def xz_compress(self, d):
        import oracle_py
        return oracle_py.xz_compress(d)


def test_reference_suite_cpu_tier(golden):
    v = next(x for x in golden.vectors() if x["name"] == "foo.txt.lzma")
    import oracle_py
    foo = oracle_py.lzma_decompress(golden.compressed(v)).out
    assert len(foo) == v["plain_len"]
    run_reference_suite(_CpuTierApi(), foo)
