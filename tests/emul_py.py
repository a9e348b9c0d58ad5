"""Loads tests/host_emulation/liblzb_emul.so: K1's decode core compiled as plain C++ (1-lane warp) behind the
product's host planning code.  CPU-tier test infrastructure only; never used by the product or by GPU tests."""
import ctypes as C
import os
import subprocess

import numpy as np

from lzma_rs_b200 import _native

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_DIR = os.path.join(_HERE, "host_emulation")
_SO = os.path.join(_DIR, "liblzb_emul.so")
_SRCS = [os.path.join(_DIR, "emul.cpp"), os.path.join(_ROOT, "lzma_rs_b200", "csrc", "lzb_plan.cpp")]
_DEPS = _SRCS + [os.path.join(_ROOT, "lzma_rs_b200", "csrc", f) for f in
                 ("lzb_decode_core.h", "lzb_plan.h", "lzb_types.h", "lzb_sched.h")] + [os.path.join(_ROOT, "include", "lzma_b200.h")]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if (not os.path.exists(_SO)) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in _DEPS):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-I" + os.path.join(_ROOT, "include"),
                                   "-I" + os.path.join(_ROOT, "lzma_rs_b200", "csrc"), "-o", _SO] + _SRCS)
        _lib = C.CDLL(_SO)
        vp = C.c_void_p
        _lib.emul_decode_batch.argtypes = [C.c_int, C.POINTER(_native.Options), vp, vp, C.c_uint32, vp, vp, vp, vp, vp]
        _lib.emul_scan_capacity.argtypes = [C.c_int, C.POINTER(_native.Options), vp, C.c_uint64]
        _lib.emul_scan_capacity.restype = C.c_uint64
        _lib.emul_sched_plan.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp]
        _lib.emul_sched_plan.restype = C.c_uint32
        _lib.emul_sched_simulate.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.emul_sched_simulate.restype = C.c_double
        _lib.emul_raw_create.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.emul_raw_create.restype = vp
        _lib.emul_raw_reset.argtypes = [vp]
        _lib.emul_raw_destroy.argtypes = [vp]
        _lib.emul_raw_decompress.argtypes = [vp, C.POINTER(_native.Options), C.c_char_p, C.c_uint64, vp, C.c_uint64, vp, vp,
                                             C.POINTER(_native.Status)]
        _lib.lzb_format_error.argtypes = [C.POINTER(_native.Status), C.c_char_p, C.c_size_t]
        _lib.lzb_format_error.restype = C.c_size_t
    return _lib


class Result:
    def __init__(self, data, consumed, status, display):
        self.data, self.consumed, self.status, self.display = data, consumed, status, display

    @property
    def ok(self):
        return int(self.status["code"]) == 0


def decode_batch(fmt, streams, opt=None, capacities=None):
    """Same contract as lzma_rs_b200.Context.decode_batch, executed by the host emulation."""
    L = lib()
    opt = opt or _native.make_options()
    blob, in_off = _native.pack_streams(streams)
    n = len(streams)
    if capacities is None:
        capacities = [L.emul_scan_capacity(fmt, C.byref(opt), blob.ctypes.data + int(in_off[i]),
                                           int(in_off[i + 1] - in_off[i])) for i in range(n)]
    capacities = np.asarray(capacities, dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((capacities + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
    out = np.zeros(int(out_off[-1]) + 16, dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    consumed = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=_native.STATUS_DTYPE)
    rc = L.emul_decode_batch(fmt, C.byref(opt), blob.ctypes.data, in_off.ctypes.data, n, out.ctypes.data,
                             out_off.ctypes.data, out_len.ctypes.data, consumed.ctypes.data, st.ctypes.data)
    assert rc == 0, rc
    res = []
    for i in range(n):
        o = int(out_off[i])
        disp = "" if st[i]["code"] == 0 else _native.format_status(L, st[i])
        res.append(Result(out[o:o + int(out_len[i])].tobytes(), int(consumed[i]), st[i].copy(), disp))
    return res


def sched_plan(work, sms=148, warps=28):
    """lzb_sched::plan on a queue of per-stream work (longest first).  Returns (order, info dict)."""
    L = lib()
    w = np.ascontiguousarray(work, dtype=np.float64)
    order = np.zeros(len(w) + sms * warps, dtype=np.uint32)
    info = np.zeros(4, dtype=np.uint32)
    model = np.zeros(2, dtype=np.float64)
    k = L.emul_sched_plan(w.ctypes.data, len(w), sms, warps, order.ctypes.data, info.ctypes.data, model.ctypes.data)
    return order[:k].copy(), dict(throttled=bool(info[0]), n_static=int(info[1]), grid=int(info[2]), parked=int(info[3]),
                                  plain=float(model[0]), predicted=float(model[1]))


def sched_simulate(work, counts=None, sms=148, warps=28):
    L = lib()
    w = np.ascontiguousarray(work, dtype=np.float64)
    if counts is None:
        return L.emul_sched_simulate(w.ctypes.data, len(w), None, 0, sms, warps)
    c = np.ascontiguousarray(counts, dtype=np.uint32)
    return L.emul_sched_simulate(w.ctypes.data, len(w), c.ctypes.data, len(c), sms, warps)


class RawHandle:
    """CPU-tier twin of lzma_rs_b200.RawHandle (lzb_raw_*): K1's CARRY path through the host emulation."""

    def __init__(self, fmt, lc, lp, pb, dict_size):
        self._h = lib().emul_raw_create(fmt, lc, lp, pb, dict_size)

    def reset(self):
        lib().emul_raw_reset(self._h)

    def decompress(self, data, options=None):
        L = lib()
        opt = options._native() if options is not None else _native.make_options()
        data = bytes(data)
        cap = max(1 << 16, len(data) * 8)
        while True:
            out = np.zeros(cap + 16, dtype=np.uint8)
            out_len, consumed = C.c_uint64(), C.c_uint64()
            st = _native.Status()
            code = L.emul_raw_decompress(self._h, C.byref(opt), data, len(data), out.ctypes.data, cap, C.byref(out_len),
                                         C.byref(consumed), C.byref(st))
            if code == -1 and cap < (1 << 30):
                cap *= 4
                continue
            break
        row = np.zeros((), dtype=_native.STATUS_DTYPE)
        row["code"], row["kind"], row["a0"], row["a1"], row["a2"] = st.code, st.kind, st.a0, st.a1, st.a2
        disp = "" if st.code == 0 else _native.format_status(L, st)
        return Result(out[:out_len.value].tobytes(), consumed.value, row, disp)

    def __del__(self):
        try:
            lib().emul_raw_destroy(self._h)
        except Exception:
            pass
