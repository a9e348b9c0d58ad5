"""ctypes binding of the C ABI (include/lzma_b200.h) exported by lzma_rs_b200/liblzma_b200.so.

There is no CPU fallback: if the CUDA library is missing or no GPU is present every decode call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# LZMA_B200_LIB: another build of the same library (kernel experiments: tools/variants.py); default = the in-tree build
LIB_PATH = os.environ.get("LZMA_B200_LIB") or os.path.join(_HERE, "liblzma_b200.so")

FMT_LZMA, FMT_LZMA2, FMT_XZ = 0, 1, 2
RC_OK, RC_BAD_ARG, RC_NO_DEVICE, RC_CUDA, RC_OOM = 0, -1, -2, -3, -4
KIND_OK, KIND_IO, KIND_HEADER_TOO_SHORT, KIND_LZMA, KIND_XZ, KIND_INTERNAL = range(6)
E_CAPACITY, E_UNSUPPORTED = -1, -2
E_IO_EOF, E_UNPACKED_MISMATCH = 1, 6  # LZB_E_IO_EOF, LZB_E_UNPACKED_MISMATCH


class Options(C.Structure):  # lzb_options
    _fields_ = [("unpacked_mode", C.c_uint8), ("has_provided", C.c_uint8), ("has_memlimit", C.c_uint8),
                ("allow_incomplete", C.c_uint8), ("reserved", C.c_uint8 * 4), ("provided", C.c_uint64), ("memlimit", C.c_uint64)]


class CompressOptions(C.Structure):  # lzb_compress_options
    _fields_ = [("skip_size_field", C.c_uint8), ("has_value", C.c_uint8), ("reserved", C.c_uint8 * 6),
                ("value", C.c_uint64)]


class IpcHandle(C.Structure):  # lzb_ipc_handle
    _fields_ = [("handle", C.c_uint8 * 64), ("offset", C.c_uint64), ("bytes", C.c_uint64)]


class Status(C.Structure):  # lzb_status
    _fields_ = [("code", C.c_int32), ("kind", C.c_int32), ("a0", C.c_uint64), ("a1", C.c_uint64), ("a2", C.c_uint64)]


STATUS_DTYPE = np.dtype([("code", "<i4"), ("kind", "<i4"), ("a0", "<u8"), ("a1", "<u8"), ("a2", "<u8")])
assert STATUS_DTYPE.itemsize == C.sizeof(Status)

# every symbol include/lzma_b200.h declares
EXPORTS = ["lzb_create", "lzb_destroy", "lzb_last_error", "lzb_abi_version", "lzb_scan", "lzb_decode_batch",
           "lzb_decode_batch_device", "lzb_batch_prepare", "lzb_batch_launch", "lzb_batch_collect", "lzb_batch_destroy",
           "lzb_batch_kernels_per_launch", "lzb_batch_kernel_name", "lzb_decompress_alloc", "lzb_free", "lzb_crc_device", "lzb_format_error",
           "lzb_encode_bound", "lzb_encode_batch", "lzb_encode_batch_device",
           "lzb_scan_device", "lzb_batch_prepare_device", "lzb_create_multi", "lzb_destroy_multi", "lzb_multi_device_count",
           "lzb_multi_ctx", "lzb_multi_last_error", "lzb_decode_batch_multi", "lzb_ipc_export", "lzb_ipc_open",
           "lzb_ipc_close", "lzb_decode_batch_peer", "lzb_raw_create", "lzb_raw_reset", "lzb_raw_decompress",
           "lzb_raw_destroy"]

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Loads liblzma_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(lzma_rs_b200 has no CPU fallback)")
    _lib = bind(LIB_PATH)
    return _lib


def bind(path):
    """dlopen one build of the library and declare the C ABI's argument types (tools/kbench.py binds several)."""
    lib = C.CDLL(path)
    vp, u64p = C.c_void_p, C.c_void_p
    lib.lzb_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.lzb_destroy.argtypes = [vp]
    lib.lzb_last_error.argtypes = [vp]
    lib.lzb_last_error.restype = C.c_char_p
    lib.lzb_scan.argtypes = [vp, C.c_int, C.POINTER(Options), vp, u64p, C.c_uint32, u64p]
    lib.lzb_decode_batch.argtypes = [vp, C.c_int, C.POINTER(Options), vp, u64p, C.c_uint32, vp, u64p, u64p, u64p, vp]
    lib.lzb_decode_batch_device.argtypes = [vp, C.c_int, C.POINTER(Options), vp, u64p, C.c_uint32, vp, u64p, u64p,
                                            u64p, vp, vp]
    lib.lzb_batch_prepare.argtypes = [vp, C.c_int, C.POINTER(Options), vp, u64p, C.c_uint32, vp, u64p, C.POINTER(vp)]
    lib.lzb_batch_launch.argtypes = [vp, vp]
    lib.lzb_batch_collect.argtypes = [vp, vp, u64p, u64p, vp]
    lib.lzb_batch_destroy.argtypes = [vp]
    lib.lzb_batch_kernels_per_launch.argtypes = [vp]
    lib.lzb_batch_kernel_name.argtypes = [vp]
    lib.lzb_batch_kernel_name.restype = C.c_char_p
    lib.lzb_decompress_alloc.argtypes = [vp, C.c_int, C.POINTER(Options), vp, C.c_size_t, C.POINTER(vp),
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(Status)]
    lib.lzb_free.argtypes = [vp]
    lib.lzb_crc_device.argtypes = [vp, vp, u64p, u64p, C.c_uint32, vp, vp, vp]
    lib.lzb_format_error.argtypes = [C.POINTER(Status), C.c_char_p, C.c_size_t]
    lib.lzb_format_error.restype = C.c_size_t
    if hasattr(lib, "lzb_encode_batch"):  # tools/kbench.py also binds older builds
        lib.lzb_encode_bound.argtypes = [C.c_int, C.POINTER(CompressOptions), C.c_uint64]
        lib.lzb_encode_bound.restype = C.c_uint64
        lib.lzb_encode_batch.argtypes = [vp, C.c_int, C.POINTER(CompressOptions), vp, u64p, C.c_uint32, vp, u64p, u64p, vp]
        lib.lzb_encode_batch_device.argtypes = [vp, C.c_int, C.POINTER(CompressOptions), vp, u64p, C.c_uint32, vp, u64p,
                                                u64p, vp, vp]
    if hasattr(lib, "lzb_decode_batch_peer"):  # ABI v2
        lib.lzb_scan_device.argtypes = [vp, C.c_int, C.POINTER(Options), vp, vp, C.c_uint32, vp, vp, C.POINTER(C.c_uint64), vp]
        lib.lzb_batch_prepare_device.argtypes = [vp, C.c_int, C.POINTER(Options), vp, vp, C.c_uint32, vp, vp, C.POINTER(vp)]
        lib.lzb_create_multi.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
        lib.lzb_destroy_multi.argtypes = [vp]
        lib.lzb_multi_device_count.argtypes = [vp]
        lib.lzb_multi_ctx.argtypes = [vp, C.c_int]
        lib.lzb_multi_ctx.restype = vp
        lib.lzb_multi_last_error.argtypes = [vp]
        lib.lzb_multi_last_error.restype = C.c_char_p
        lib.lzb_decode_batch_multi.argtypes = [vp, C.c_int, C.POINTER(Options), vp, u64p, C.c_uint32, vp, u64p, u64p, u64p, vp, vp]
        lib.lzb_ipc_export.argtypes = [vp, vp, C.c_uint64, C.POINTER(IpcHandle)]
        lib.lzb_ipc_open.argtypes = [vp, C.POINTER(IpcHandle), C.POINTER(vp)]
        lib.lzb_ipc_close.argtypes = [vp, vp]
        lib.lzb_decode_batch_peer.argtypes = [vp, C.c_int, C.POINTER(Options), vp, u64p, C.c_uint32, vp, u64p, u64p, u64p, vp]
    if hasattr(lib, "lzb_raw_create"):
        lib.lzb_raw_create.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
        lib.lzb_raw_reset.argtypes = [vp]
        lib.lzb_raw_decompress.argtypes = [vp, C.POINTER(Options), vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t),
                                           C.POINTER(C.c_size_t), C.POINTER(Status)]
        lib.lzb_raw_destroy.argtypes = [vp]
    return lib


def format_status(lib, st_row):
    """Display string of one status record (numpy STATUS_DTYPE row or ctypes Status)."""
    if isinstance(st_row, Status):
        s = st_row
    else:
        s = Status(int(st_row["code"]), int(st_row["kind"]), int(st_row["a0"]), int(st_row["a1"]), int(st_row["a2"]))
    buf = C.create_string_buffer(512)
    lib.lzb_format_error(C.byref(s), buf, 512)
    return buf.value.decode()


def make_options(unpacked_mode=0, provided=None, memlimit=None, allow_incomplete=False):
    o = Options()
    o.allow_incomplete = 1 if allow_incomplete else 0
    o.unpacked_mode = unpacked_mode
    o.has_provided = 0 if provided is None else 1
    o.provided = 0 if provided is None else provided
    o.has_memlimit = 0 if memlimit is None else 1
    o.memlimit = 0 if memlimit is None else memlimit
    return o


def pack_streams(streams, align=16):
    """Concatenates byte strings into one uint8 blob (each start `align`-aligned).  Returns (blob, in_off[n+1])
    where stream i = blob[in_off[i]: in_off[i] + len_i]; because lzb_* take in_off[i+1] as the END of stream i, the
    packing is tight (align=1) unless the caller handles lengths itself."""
    n = len(streams)
    lens = np.fromiter((len(s) for s in streams), dtype=np.uint64, count=n)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    blob = np.empty(int(off[-1]) + 16, dtype=np.uint8)
    blob[int(off[-1]):] = 0
    for i, s in enumerate(streams):
        if len(s):
            blob[int(off[i]):int(off[i + 1])] = np.frombuffer(s, dtype=np.uint8)
    return blob, off
