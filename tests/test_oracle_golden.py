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
