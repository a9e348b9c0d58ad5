"""ctypes binding of the CPU oracle (oracle/liblzma_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (lzma_rs_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(ORACLE_DIR, "liblzma_oracle.so")

KIND_NAMES = {0: "Ok", 1: "IoError", 2: "HeaderTooShort", 3: "LzmaError", 4: "XzError"}


class _Err(C.Structure):
    _fields_ = [("kind", C.c_int), ("msg", C.c_char * 320)]


class _Opt(C.Structure):
    _fields_ = [("unpacked_mode", C.c_int), ("has_provided", C.c_int), ("provided", C.c_uint64),
                ("has_memlimit", C.c_int), ("memlimit", C.c_uint64)]


class _COpt(C.Structure):  # lzo_compress_options
    _fields_ = [("skip_size_field", C.c_int), ("has_value", C.c_int), ("value", C.c_uint64)]


class _Res(C.Structure):
    _fields_ = [("out", C.POINTER(C.c_uint8)), ("out_len", C.c_size_t), ("consumed", C.c_size_t), ("err", _Err)]


def build():
    """Compile the oracle if the .so is missing or stale."""
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("lzma_oracle.c", "lzma_oracle_enc.c", "lzma_oracle.h")]
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.lzo_lzma_decompress.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Opt), C.POINTER(_Res)]
        _lib.lzo_lzma2_decompress.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Res)]
        _lib.lzo_xz_decompress.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Res)]
        _lib.lzo_result_free.argtypes = [C.POINTER(_Res)]
        _lib.lzo_error_display.argtypes = [C.POINTER(_Err), C.c_char_p, C.c_size_t]
        _lib.lzo_decompress_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_int]
        for f in (_lib.lzo_lzma2_compress, _lib.lzo_xz_compress):
            f.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
        _lib.lzo_lzma_compress.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_COpt), C.POINTER(C.POINTER(C.c_uint8)),
                                           C.POINTER(C.c_size_t)]
        _lib.lzo_raw_new.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, C.c_int,
                                     C.c_uint64]
        _lib.lzo_raw_new.restype = C.c_void_p
        _lib.lzo_raw_reset.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64]
        _lib.lzo_raw_decompress.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(_Res)]
        _lib.lzo_raw_free.argtypes = [C.c_void_p]
        _lib.lzo_buffer_free.argtypes = [C.POINTER(C.c_uint8)]
        _lib.lzo_crc32.argtypes = [C.c_char_p, C.c_size_t]
        _lib.lzo_crc32.restype = C.c_uint32
        _lib.lzo_crc64.argtypes = [C.c_char_p, C.c_size_t]
        _lib.lzo_crc64.restype = C.c_uint64
    return _lib


class OracleResult:
    """(output bytes handed to the sink, consumed input bytes, error kind, Display string)."""

    def __init__(self, out, consumed, kind, msg, display):
        self.out, self.consumed, self.kind, self.msg, self.display = out, consumed, kind, msg, display

    @property
    def ok(self):
        return self.kind == 0

    def __repr__(self):
        return f"OracleResult(len={len(self.out)}, consumed={self.consumed}, kind={KIND_NAMES[self.kind]}, {self.display!r})"


def _finish(res):
    out = C.string_at(res.out, res.out_len) if res.out_len else b""
    buf = C.create_string_buffer(400)
    lib().lzo_error_display(C.byref(res.err), buf, 400)
    r = OracleResult(out, res.consumed, res.err.kind, res.err.msg.decode(), buf.value.decode())
    lib().lzo_result_free(C.byref(res))
    return r


def make_options(unpacked_mode=0, provided=None, memlimit=None):
    """unpacked_mode: 0 ReadFromHeader, 1 ReadHeaderButUseProvided(provided), 2 UseProvided(provided)."""
    o = _Opt()
    o.unpacked_mode = unpacked_mode
    o.has_provided = 0 if provided is None else 1
    o.provided = 0 if provided is None else provided
    o.has_memlimit = 0 if memlimit is None else 1
    o.memlimit = 0 if memlimit is None else memlimit
    return o


def lzma_decompress(data, unpacked_mode=0, provided=None, memlimit=None):
    res = _Res()
    opt = make_options(unpacked_mode, provided, memlimit)
    lib().lzo_lzma_decompress(bytes(data), len(data), C.byref(opt), C.byref(res))
    return _finish(res)


def lzma2_decompress(data):
    res = _Res()
    lib().lzo_lzma2_decompress(bytes(data), len(data), C.byref(res))
    return _finish(res)


def xz_decompress(data):
    res = _Res()
    lib().lzo_xz_decompress(bytes(data), len(data), C.byref(res))
    return _finish(res)


class RawDecoder:
    """decompress::raw::LzmaDecoder (fmt 0) / Lzma2Decoder (fmt 1) restated in the oracle: the DecoderState survives
    from one decompress() to the next until reset() (lzma.rs:597-648, lzma2.rs:11-82)."""

    def __init__(self, fmt, lc=0, lp=0, pb=0, dict_size=0, unpacked=None, memlimit=None):
        self._h = lib().lzo_raw_new(fmt, lc, lp, pb, dict_size, 0 if unpacked is None else 1, unpacked or 0,
                                    0 if memlimit is None else 1, memlimit or 0)

    def reset(self, unpacked=...):
        if unpacked is ...:
            lib().lzo_raw_reset(self._h, 0, 0, 0)
        else:
            lib().lzo_raw_reset(self._h, 1, 0 if unpacked is None else 1, unpacked or 0)

    def decompress(self, data):
        res = _Res()
        lib().lzo_raw_decompress(self._h, bytes(data), len(data), C.byref(res))
        return _finish(res)

    def __del__(self):
        try:
            lib().lzo_raw_free(self._h)
        except Exception:
            pass


def crc32(data):
    return lib().lzo_crc32(bytes(data), len(data))


def crc64(data):
    return lib().lzo_crc64(bytes(data), len(data))


def decompress_batch(fmt, blob, in_off, out_off, nthreads):
    """CPU-baseline batch driver.  blob: np.uint8 array; in_off/out_off: np.uint64 (n+1).
    Returns (out np.uint8, out_len np.uint64, kinds np.int32, failed)."""
    n = len(in_off) - 1
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    out = np.empty(int(out_off[-1]), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    kinds = np.zeros(n, dtype=np.int32)
    failed = lib().lzo_decompress_batch(fmt, blob.ctypes.data, in_off.ctypes.data, n, out.ctypes.data,
                                        out_off.ctypes.data, out_len.ctypes.data, kinds.ctypes.data, nthreads)
    return out, out_len, kinds, failed


def _take(fn, *args):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    assert fn(*args, C.byref(out), C.byref(n)) == 0
    data = C.string_at(out, n.value) if n.value else b""
    lib().lzo_buffer_free(out)
    return data


def lzma_compress(data, skip_size_field=False, value=None):
    """lzma_compress_with_options (src/lib.rs:72-80): WriteToHeader(value) / SkipWritingToHeader."""
    o = _COpt(1 if skip_size_field else 0, 0 if value is None else 1, 0 if value is None else value)
    return _take(lib().lzo_lzma_compress, bytes(data), len(data), C.byref(o))


def lzma2_compress(data):
    return _take(lib().lzo_lzma2_compress, bytes(data), len(data))


def xz_compress(data):
    return _take(lib().lzo_xz_compress, bytes(data), len(data))


class Stream:
    """decompress::Stream restated in the oracle (lzo_stream_*): write / write_all / finish like the reference's."""

    def __init__(self, unpacked_mode=0, provided=None, memlimit=None, allow_incomplete=False):
        L = lib()
        L.lzo_stream_new.restype = C.c_void_p
        L.lzo_stream_new.argtypes = [C.POINTER(_Opt), C.c_int]
        L.lzo_stream_write.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(_Err)]
        L.lzo_stream_finish.argtypes = [C.c_void_p, C.POINTER(_Res)]
        opt = make_options(unpacked_mode, provided, memlimit)
        self._h = L.lzo_stream_new(C.byref(opt), 1 if allow_incomplete else 0)

    def write(self, data):
        """Returns (consumed, None) or (0, OracleResult-like error with .kind/.display)."""
        n, err = C.c_size_t(), _Err()
        kind = lib().lzo_stream_write(self._h, bytes(data), len(data), C.byref(n), C.byref(err))
        if kind:
            buf = C.create_string_buffer(400)
            lib().lzo_error_display(C.byref(err), buf, 400)
            return 0, OracleResult(b"", 0, kind, err.msg.decode(), buf.value.decode())
        return n.value, None

    def write_all(self, data):
        """io::Write::write_all: None on success, else the error (WriteZero when a write accepts nothing)."""
        data = bytes(data)
        while data:
            k, err = self.write(data)
            if err is not None:
                return err
            if k == 0:
                return OracleResult(b"", 0, 1, "failed to write whole buffer", "io error: failed to write whole buffer")
            data = data[k:]
        return None

    def finish(self):
        res = _Res()
        lib().lzo_stream_finish(self._h, C.byref(res))
        self._h = None
        return _finish(res)
