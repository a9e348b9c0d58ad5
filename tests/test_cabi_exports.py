"""CPU tier: the C-ABI library loads, exports every symbol include/lzma_b200.h declares, refuses to work
without a GPU (no CPU fallback), and renders the reference's error strings."""
import ctypes as C
import os
import re

import pytest

from lzma_rs_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lzma_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lzb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.lzb_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _native.load()
    h = C.c_void_p()
    assert lib.lzb_create(C.byref(h), -1) == _native.RC_NO_DEVICE
    import lzma_rs_b200 as L
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.lzma2_decompress(b"\0")


@pytest.mark.parametrize("code,kind,a,want", [
    (1, 1, (0, 0, 0), "io error: failed to fill whole buffer"),
    (2, 2, (0, 0, 0), "header too short: failed to fill whole buffer"),
    (3, 3, (255, 0, 0), "lzma error: LZMA header invalid properties: 255 must be < 225"),
    (6, 3, (10, 12, 0), "lzma error: Expected unpacked size of 10 but decompressed to 12"),
    (10, 3, (3226, 2, 0), "lzma error: LZ distance 3226 is beyond output size 2"),
    (11, 3, (0, 0, 0), "lzma error: exceeded memory limit of 0"),
    (13, 3, (5, 0, 0), "lzma error: LZMA2 invalid status 5, must be 0, 1, 2 or >= 128"),
    (18, 3, (3, 2, 0), "lzma error: LZMA2 invalid properties: lc + lp (3 + 2) must be <= 4"),
    (54, 4, (0x01234567, 0x8B0D303E, 0), "xz error: Invalid footer CRC32: expected 0x01234567 but got 0x8b0d303e"),
    (44, 4, (1, 2, 0), "xz error: Invalid block CRC64, expected 0x0000000000000001 but got 0x0000000000000002"),
    (53, 4, (1, 4, 0), "xz error: Flags in header (StreamFlags { check_method: Crc32 }) does not match footer "
                       "(StreamFlags { check_method: Crc64 })"),
    (47, 4, (0, 100, 104), "xz error: Invalid index for record 0: unpadded size (100) does not match index (104)"),
])
def test_error_display_strings(code, kind, a, want):
    st = _native.Status(code, kind, *a)
    assert _native.format_status(_native.load(), st) == want
