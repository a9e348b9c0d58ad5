"""Pins the CPU oracle against every golden vector the reference's own tests hold for the path
(SURVEY.md 8(c)): fixtures + inline KATs + expected error strings."""
import hashlib

import pytest

import oracle_py as oracle

DECODERS = {"lzma": oracle.lzma_decompress, "xz": oracle.xz_decompress}


def _check_error(v, display):
    how, want = v["error_match"], v["error"]
    if how == "exact":
        assert display == want
    elif how == "prefix":
        assert display.startswith(want)
    else:
        assert want in display


def test_golden_vectors_decode(golden):
    n = 0
    for v in golden.vectors(errors=False):
        comp = golden.compressed(v)
        r = DECODERS[v["format"]](comp)
        assert r.ok, (v["name"], r.display)
        assert len(r.out) == v["plain_len"], v["name"]
        assert hashlib.sha256(r.out).hexdigest() == v["plain_sha256"], v["name"]
        assert r.consumed == len(comp), v["name"]
        n += 1
    assert n == 18


def test_golden_error_strings(golden):
    n = 0
    for v in golden.vectors(errors=True):
        r = DECODERS[v["format"]](golden.compressed(v))
        assert not r.ok, v["name"]
        _check_error(v, r.display)
        # partial output: whatever reached the sink before the error (whole validated xz blocks)
        assert hashlib.sha256(r.out).hexdigest() == v["plain_sha256"], v["name"]
        n += 1
    assert n == 4


def test_crc_check_values():
    # crc crate catalogue check values for CRC_32_ISO_HDLC / CRC_64_XZ (src/xz/crc.rs:3-4)
    assert oracle.crc32(b"123456789") == 0xCBF43926
    assert oracle.crc64(b"123456789") == 0x995DC9BBDF1939FA


@pytest.mark.parametrize("mode,provided,ok", [
    (0, None, True),       # ReadFromHeader (header says unknown -> end marker)
    (1, 12, True),         # ReadHeaderButUseProvided(Some(12))
    (1, None, True),       # ReadHeaderButUseProvided(None)
    (1, 5, False),         # provided too small: a literal run stops exactly at 5 -> OK? (see below)
])
def test_unpacked_size_modes_on_inline_hello(golden, mode, provided, ok):
    # tests/lzma.rs:237-303 exercise the three UnpackedSize modes; here on the inline hello vector.
    v = next(x for x in golden.vectors() if x["name"] == "inline-hello.lzma")
    r = oracle.lzma_decompress(golden.compressed(v), unpacked_mode=mode, provided=provided)
    if provided == 5:
        # stops as soon as len >= 5 (lzma.rs:442-445): literals only, so exactly 5 bytes, no error
        assert r.ok and r.out == b"Hello"
    else:
        assert r.ok == ok and r.out == b"Hello world\n"


def test_use_provided_skips_size_field(golden):
    v = next(x for x in golden.vectors() if x["name"] == "inline-hello.lzma")
    comp = golden.compressed(v)
    stripped = comp[:5] + comp[13:]  # drop the 8-byte size field (UseProvided, options.rs:38-42)
    r = oracle.lzma_decompress(stripped, unpacked_mode=2, provided=12)
    assert r.ok and r.out == b"Hello world\n"
    r = oracle.lzma_decompress(stripped, unpacked_mode=2, provided=None)
    assert r.ok and r.out == b"Hello world\n"


def test_memlimit_error(golden):
    # tests/lzma.rs:306-336: memlimit Some(0) -> "exceeded memory limit of 0"
    v = next(x for x in golden.vectors() if x["name"] == "inline-hello.lzma")
    r = oracle.lzma_decompress(golden.compressed(v), memlimit=0)
    assert not r.ok and "exceeded memory limit of 0" in r.display and r.out == b""


def test_oracle_accepts_what_liblzma_accepts():
    """The reference's own differential oracle is liblzma (tests/lzma.rs:109-114, fuzz/fuzz_targets/compare_xz.rs).  lzma-rs
    is the more lenient of the two except for features it lacks, so over mutated and field-fuzzed streams: whatever
    liblzma decodes, the oracle decodes to the same bytes -- apart from check ids other than None/CRC32/CRC64/SHA-256,
    which the reference rejects (src/xz/mod.rs:65-75) and liblzma skips."""
    import lzma
    import os
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import corpus
    import fuzz_soak
    import oracle_py as oracle

    def lib(fmt, b):
        try:
            d = lzma.LZMADecompressor(format=lzma.FORMAT_XZ if fmt == 2 else lzma.FORMAT_ALONE)
            out = d.decompress(b)
            return out if d.eof and not d.unused_data else None
        except Exception:
            return None

    rnd = random.Random(4)
    accepted = 0
    for _ in range(4):
        seeds = fuzz_soak.seeds_for(rnd, corpus)
        items = [(2, fuzz_soak.xz_structured(rnd, corpus)) for _ in range(150)]
        items += [(f, fuzz_soak.mutate(rnd, seeds[f][i % len(seeds[f])])) for f in (0, 2) for i in range(150)]
        for fmt, b in items:
            ref = lib(fmt, b)
            if ref is None:
                continue
            r = oracle.xz_decompress(b) if fmt == 2 else oracle.lzma_decompress(b)
            if not r.ok and "Invalid check method" in r.display:
                continue
            accepted += 1
            assert r.ok and r.out == ref, (fmt, r.display, b[:24].hex())
    assert accepted > 50
