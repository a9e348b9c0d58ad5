"""lzma_rs_b200 -- B200-native many-stream LZMA / LZMA2 / XZ decompressor (host-side mirror of the lzma-rs API).

Mirrors the decode entry points of gendx/lzma-rs (src/lib.rs:44-60, 83-88, 100-105) over the C ABI in
include/lzma_b200.h:

    lzma_decompress(input, output)                        -> lzma_rs::lzma_decompress
    lzma_decompress_with_options(input, output, options)  -> lzma_rs::lzma_decompress_with_options
    lzma2_decompress(input, output)                       -> lzma_rs::lzma2_decompress
    xz_decompress(input, output)                          -> lzma_rs::xz_decompress

`input` is a bytes-like object or a binary reader (`.read()`), `output` a binary writer (`.write()`), the same
roles io::BufRead / io::Write play in the reference.  Errors are raised as `error.Error` subclasses whose `str()`
is the reference's Display string.  On error the reference's partial output has already been written to `output`.

The `*_batch` functions are the many-stream form the GPU is built for.  Every byte is decoded by the CUDA kernels
in liblzma_b200.so -- there is no CPU fallback; without the library or a GPU the calls raise.
"""
import ctypes as _C
import enum as _enum

import numpy as _np

from . import _native

__all__ = ["lzma_decompress", "lzma_decompress_with_options", "lzma2_decompress", "xz_decompress",
           "lzma_decompress_batch", "lzma2_decompress_batch", "xz_decompress_batch", "decompress", "error", "Context",
           "Stream", "StreamResult", "compress", "lzma_compress", "lzma_compress_with_options", "lzma2_compress",
           "xz_compress", "lzma_compress_batch", "lzma2_compress_batch", "xz_compress_batch"]


class error:  # namespace mirroring lzma_rs::error (src/error.rs:7-37)
    class Error(Exception):
        """lzma_rs::error::Error"""
        prefix = ""

        def __init__(self, display, status=None, partial_output=b""):
            super().__init__(display)
            self.status = status
            self.partial_output = partial_output

    class IoError(Error):
        pass

    class HeaderTooShort(Error):
        pass

    class LzmaError(Error):
        pass

    class XzError(Error):
        pass

    class InternalError(Error):
        """Not a reference error: capacity/limits of the GPU path or a CUDA failure."""


_KIND_TO_EXC = {_native.KIND_IO: error.IoError, _native.KIND_HEADER_TOO_SHORT: error.HeaderTooShort,
                _native.KIND_LZMA: error.LzmaError, _native.KIND_XZ: error.XzError,
                _native.KIND_INTERNAL: error.InternalError}


class decompress:  # namespace mirroring lzma_rs::decompress (src/decode/options.rs:3-43)
    class UnpackedSizeMode(_enum.IntEnum):
        ReadFromHeader = 0
        ReadHeaderButUseProvided = 1
        UseProvided = 2

    class UnpackedSize:
        """UnpackedSize::{ReadFromHeader, ReadHeaderButUseProvided(Option<u64>), UseProvided(Option<u64>)}"""

        def __init__(self, mode=0, value=None):
            self.mode, self.value = int(mode), value

        @classmethod
        def ReadFromHeader(cls):
            return cls(0)

        @classmethod
        def ReadHeaderButUseProvided(cls, x):
            return cls(1, x)

        @classmethod
        def UseProvided(cls, x):
            return cls(2, x)

    class Options:
        """decompress::Options { unpacked_size, memlimit, allow_incomplete (stream API only) }"""

        def __init__(self, unpacked_size=None, memlimit=None, allow_incomplete=False):
            self.unpacked_size = unpacked_size or decompress.UnpackedSize.ReadFromHeader()
            self.memlimit = memlimit
            self.allow_incomplete = allow_incomplete

        def _native(self):
            return _native.make_options(self.unpacked_size.mode, self.unpacked_size.value, self.memlimit, self.allow_incomplete)

    class raw:  # namespace mirroring lzma_rs::decompress::raw (feature `raw_decoder`, src/lib.rs:29-35)
        """Reusable raw decoders.  Like the reference's, a decoder keeps its probability state between two `decompress`
        calls unless `reset` is called in between (the state lives in device memory: lzb_raw_* of the C ABI)."""

        class LzmaProperties:
            """LzmaProperties { lc, lp, pb } (src/decode/lzma.rs:41-66)."""

            def __init__(self, lc, lp, pb):
                self.lc, self.lp, self.pb = int(lc), int(lp), int(pb)

            def validate(self):
                assert self.lc <= 8 and self.lp <= 4 and self.pb <= 4  # lzma.rs:62-66 (the reference panics)

        class LzmaParams:
            """LzmaParams::new(properties, dict_size, unpacked_size) (src/decode/lzma.rs:68-93)."""

            def __init__(self, properties, dict_size, unpacked_size=None):
                self.properties, self.dict_size, self.unpacked_size = properties, int(dict_size), unpacked_size

            @classmethod
            def read_header(cls, input, options=None):
                """LzmaParams::read_header (lzma.rs:96-161): consumes the 13-byte (5-byte with UseProvided) header."""
                options = options or decompress.Options()
                us = options.unpacked_size
                head = input.read(1)
                if len(head) < 1:
                    raise error.HeaderTooShort("header too short: failed to fill whole buffer")
                props = head[0]
                if props >= 225:
                    raise error.LzmaError(f"lzma error: LZMA header invalid properties: {props} must be < 225")
                lc, lp, pb = props % 9, (props // 9) % 5, props // 45
                d = input.read(4)
                if len(d) < 4:
                    raise error.HeaderTooShort("header too short: failed to fill whole buffer")
                dict_size = max(int.from_bytes(d, "little"), 0x1000)
                if us.mode == decompress.UnpackedSizeMode.UseProvided:
                    size = us.value
                else:
                    h = input.read(8)
                    if len(h) < 8:
                        raise error.HeaderTooShort("header too short: failed to fill whole buffer")
                    hv = int.from_bytes(h, "little")
                    if us.mode == decompress.UnpackedSizeMode.ReadHeaderButUseProvided:
                        size = us.value
                    else:
                        size = None if hv == 0xFFFFFFFFFFFFFFFF else hv
                return cls(decompress.raw.LzmaProperties(lc, lp, pb), dict_size, size)

        class LzmaDecoder:
            """raw::LzmaDecoder (src/decode/lzma.rs:597-648): `input` is a headerless LZMA stream.  Like the reference's,
            the decoder keeps its DecoderState (probabilities, state, rep distances) from one `decompress` to the next
            until `reset()`; every call starts an empty output window (lzb_raw_* in the C ABI)."""

            def __init__(self, params, memlimit=None, ctx=None):
                params.properties.validate()
                self.params, self.memlimit, self._ctx = params, memlimit, ctx
                self._unpacked = params.unpacked_size
                self._raw = None

            def _handle(self):
                if self._raw is None:
                    pr = self.params.properties
                    self._raw = (self._ctx or _ctx()).raw_new(_native.FMT_LZMA, pr.lc, pr.lp, pr.pb, self.params.dict_size)
                return self._raw

            def reset(self, unpacked_size=...):
                """reset(None) keeps the size; reset(Some(x)) replaces it (lzma.rs:620-627): pass nothing or x."""
                if unpacked_size is not ...:
                    self._unpacked = unpacked_size
                if self._raw is not None:
                    self._raw.reset()

            def decompress(self, input, output=None):
                opts = decompress.Options(decompress.UnpackedSize.UseProvided(self._unpacked), self.memlimit)
                data, reader = _read_all(input)
                r = self._handle().decompress(data, opts)
                return _deliver(r, len(data), 0, reader, output)

        class Lzma2Decoder:
            """raw::Lzma2Decoder (src/decode/lzma2.rs:11-82); state kept between calls like the reference's."""

            def __init__(self, ctx=None):
                self._ctx, self._raw = ctx, None

            def _handle(self):
                if self._raw is None:
                    self._raw = (self._ctx or _ctx()).raw_new(_native.FMT_LZMA2, 0, 0, 0, 0)
                return self._raw

            def reset(self):
                if self._raw is not None:
                    self._raw.reset()

            def decompress(self, input, output=None):
                data, reader = _read_all(input)
                r = self._handle().decompress(data, None)
                return _deliver(r, len(data), 0, reader, output)


class compress:  # namespace mirroring lzma_rs::compress (src/encode/options.rs:1-30)
    class UnpackedSize:
        """UnpackedSize::{WriteToHeader(Option<u64>), SkipWritingToHeader}"""

        def __init__(self, skip=False, value=None):
            self.skip, self.value = skip, value

        @classmethod
        def WriteToHeader(cls, x=None):
            return cls(False, x)

        @classmethod
        def SkipWritingToHeader(cls):
            return cls(True)

    class Options:
        """compress::Options { unpacked_size } -- default WriteToHeader(None): unknown size + end marker"""

        def __init__(self, unpacked_size=None):
            self.unpacked_size = unpacked_size or compress.UnpackedSize.WriteToHeader(None)

        def _native(self):
            u = self.unpacked_size
            return _native.CompressOptions(1 if u.skip else 0, 0 if (u.skip or u.value is None) else 1, (_C.c_uint8 * 6)(),
                                           0 if (u.skip or u.value is None) else u.value)


class Stream:
    """lzma_rs::decompress::Stream (feature `stream`, src/decode/stream.rs:66-346) as a façade over the batch path.

    The reference decodes incrementally inside `write`; the GPU path decodes whole streams, so this class buffers what
    is written and decodes in `finish()`.  Same results, later errors: header errors are still raised by `write`
    (stream.rs:157-190 parses the header as soon as its bytes are there), data errors surface in `finish()` instead of
    in the `write` that delivered the bad bytes.  `Options.allow_incomplete` is honoured (stream.rs:136-147)."""

    def __init__(self, output, options=None, ctx=None):
        self._out, self._opt, self._ctx = output, options or decompress.Options(), ctx
        self._buf, self._failed, self._done = bytearray(), False, False

    @classmethod
    def new_with_options(cls, options, output):
        return cls(output, options)

    def get_output(self):
        return None if self._failed or self._done else self._out

    def write(self, data):
        if self._failed or self._done:
            return 0  # stream.rs:230,310: the state is gone after a failed write; nothing is consumed
        first = not self._buf
        self._buf += data
        if first and self._buf and self._buf[0] >= 225:  # LzmaParams::read_header, lzma.rs:104-109
            self._failed = True
            raise error.LzmaError(f"lzma error: LZMA header invalid properties: {self._buf[0]} must be < 225")
        return len(data)

    write_all = write

    def flush(self):
        if not (self._failed or self._done) and hasattr(self._out, "flush"):
            self._out.flush()

    def finish(self):
        """Consumes the stream and returns the output sink (stream.rs:119-151)."""
        if self._failed or self._done:
            raise error.LzmaError("lzma error: can't finish stream because of previous write error")
        self._done = True
        if not self._buf:
            return self._out
        hdr = 5 if self._opt.unpacked_size.mode == decompress.UnpackedSizeMode.UseProvided else 13
        if len(self._buf) < hdr + 5:  # header + the 5 range-coder start bytes were never complete (stream.rs:123-129)
            raise error.LzmaError("lzma error: failed to read header")
        ctx, buf = (self._ctx or _ctx()), bytes(self._buf)
        if self._opt.allow_incomplete and len(buf) == hdr + 5:
            return self._out  # header and start bytes only: the reference has not decoded anything yet (stream.rs:296-305)
        strict = decompress.Options(self._opt.unpacked_size, self._opt.memlimit, False)
        r = ctx.decompress_one(_native.FMT_LZMA, buf, strict)
        data = r.data
        if self._opt.allow_incomplete and (r.ok or int(r.status["code"]) in (_native.E_IO_EOF, _native.E_UNPACKED_MISMATCH)):
            data = self._incomplete_output(ctx, buf)
            r = None
        if data:
            self._out.write(data)
        if r is not None:
            r.raise_for_status()
        if hasattr(self._out, "flush"):
            self._out.flush()
        return self._out

    def _incomplete_output(self, ctx, buf):
        """Options::allow_incomplete: what the reference's incremental decoder has produced when finish() skips the final
        pass (stream.rs:136-147) -- for input that ends inside a symbol, and also for complete unknown-size streams.  It stops at the first symbol boundary at which its input is exhausted (lzma.rs:450-452), or in
        front of the symbol it cannot complete; the batch decoder (C ABI flag `allow_incomplete`) also decodes the
        symbols behind that boundary that happen to need no further input byte.  Those are trimmed here with two more
        decodes: without the last byte the decoder stops in front of the symbol that consumes it; a decode with the
        size fixed just behind that point then ends exactly behind that symbol."""
        lenient = decompress.Options(self._opt.unpacked_size, self._opt.memlimit, True)
        full = ctx.decompress_one(_native.FMT_LZMA, buf, lenient)  # every byte of every complete symbol
        us = self._opt.unpacked_size
        known = (us.value is not None) if us.mode != decompress.UnpackedSizeMode.ReadFromHeader else \
            buf[5:13] != b"\xff" * 8
        if known:  # with a known size the reference has no "input exhausted" stop (lzma.rs:442-445 comes first)
            return full.data
        short = ctx.decompress_one(_native.FMT_LZMA, buf[:-1], lenient)
        before = len(short.data) if short.ok else 0  # output in front of the symbol that consumes the last byte
        if before >= len(full.data):
            return full.data  # that symbol is the unfinished one
        use_provided = self._opt.unpacked_size.mode == decompress.UnpackedSizeMode.UseProvided
        sized = decompress.Options(decompress.UnpackedSize(2 if use_provided else 1, before + 1), self._opt.memlimit, False)
        r = ctx.decompress_one(_native.FMT_LZMA, buf, sized)
        if r.ok:
            end = before + 1
        elif int(r.status["code"]) == _native.E_UNPACKED_MISMATCH:
            end = int(r.status["a1"])  # the symbol overshot the size: a1 = the length it reached
        else:
            end = before
        return full.data[:min(end, len(full.data))]


decompress.Stream = Stream


class StreamResult:
    """Outcome of one stream of a batch call."""
    __slots__ = ("data", "consumed", "status", "display")

    def __init__(self, data, consumed, status, display):
        self.data, self.consumed, self.status, self.display = data, consumed, status, display

    @property
    def ok(self):
        return int(self.status["code"]) == 0

    def raise_for_status(self):
        if not self.ok:
            raise _KIND_TO_EXC.get(int(self.status["kind"]), error.InternalError)(self.display, self.status, self.data)


class Context:
    """One decoder context per CUDA device (lzb_create / lzb_destroy)."""

    def __init__(self, device=-1):
        self._lib = _native.load()
        h = _C.c_void_p()
        rc = self._lib.lzb_create(_C.byref(h), device)
        if rc != _native.RC_OK:
            raise RuntimeError(f"lzb_create failed (rc={rc}): no usable CUDA device -- lzma_rs_b200 has no CPU fallback")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lzb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def last_error(self):
        return self._lib.lzb_last_error(self._h).decode()

    def scan(self, fmt, blob, in_off, options=None):
        n = len(in_off) - 1
        cap = _np.zeros(n, dtype=_np.uint64)
        opt = (options or decompress.Options())._native()
        rc = self._lib.lzb_scan(self._h, fmt, _C.byref(opt), blob.ctypes.data, in_off.ctypes.data, n, cap.ctypes.data)
        if rc != _native.RC_OK:
            raise RuntimeError(f"lzb_scan failed rc={rc}: {self.last_error()}")
        return cap

    def decode_batch(self, fmt, streams, options=None, capacities=None, retry=True):
        """streams: list of bytes-like.  Returns a list of StreamResult (same order).
        retry: streams whose size cannot be known up front (end-marker .lzma, malformed framing) and that report
        LZB_E_CAPACITY are decoded again with a larger buffer, like lzb_decompress_alloc does."""
        res = self._decode_batch_once(fmt, streams, options, capacities)
        if retry:
            for i, r in enumerate(res):
                cap = None
                while int(r.status["code"]) == _native.E_CAPACITY and (cap or 0) < 0xF0000000:
                    cap = max(int(r.status["a0"]) * 2, 1 << 16) if cap is None else cap * 2
                    r = self._decode_batch_once(fmt, [streams[i]], options, [cap])[0]
                res[i] = r
        return res

    def _decode_batch_once(self, fmt, streams, options=None, capacities=None):
        blob, in_off = _native.pack_streams(streams)
        n = len(streams)
        opt = (options or decompress.Options())._native()
        if capacities is None:
            capacities = self.scan(fmt, blob, in_off, options)
        capacities = _np.asarray(capacities, dtype=_np.uint64)
        out_off = _np.zeros(n + 1, dtype=_np.uint64)
        _np.cumsum((capacities + _np.uint64(15)) // _np.uint64(16) * _np.uint64(16), out=out_off[1:])
        out = _np.empty(int(out_off[-1]) + 16, dtype=_np.uint8)
        out_len = _np.zeros(n, dtype=_np.uint64)
        consumed = _np.zeros(n, dtype=_np.uint64)
        st = _np.zeros(n, dtype=_native.STATUS_DTYPE)
        self._decode_call(fmt, opt, blob, in_off, n, out, out_off, out_len, consumed, st)
        res = []
        for i in range(n):
            o = int(out_off[i])
            data = out[o:o + int(out_len[i])].tobytes()
            disp = "" if st[i]["code"] == 0 else _native.format_status(self._lib, st[i])
            res.append(StreamResult(data, int(consumed[i]), st[i].copy(), disp))
        return res

    def _decode_call(self, fmt, opt, blob, in_off, n, out, out_off, out_len, consumed, st):
        rc = self._lib.lzb_decode_batch(self._h, fmt, _C.byref(opt), blob.ctypes.data, in_off.ctypes.data, n,
                                        out.ctypes.data, out_off.ctypes.data, out_len.ctypes.data,
                                        consumed.ctypes.data, st.ctypes.data)
        if rc != _native.RC_OK:
            raise RuntimeError(f"lzb_decode_batch failed rc={rc}: {self.last_error()}")

    def encode_batch(self, fmt, datas, options=None):
        """lzb_encode_batch: the reference's encoders over a batch (fmt 0: literal-only .lzma, 1: stored-chunk LZMA2,
        2: stored-chunk .xz).  Returns a list of bytes.  A .lzma stream that outgrows the default bound is retried with
        the size the kernel reported."""
        n = len(datas)
        blob, in_off = _native.pack_streams(datas)
        opt = (options or compress.Options())._native()
        caps = _np.array([self._lib.lzb_encode_bound(fmt, _C.byref(opt), len(d)) for d in datas], dtype=_np.uint64)
        while True:
            out_off = _np.zeros(n + 1, dtype=_np.uint64)
            _np.cumsum((caps + _np.uint64(15)) // _np.uint64(16) * _np.uint64(16), out=out_off[1:])
            out = _np.empty(int(out_off[-1]) + 16, dtype=_np.uint8)
            out_len = _np.zeros(n, dtype=_np.uint64)
            st = _np.zeros(n, dtype=_native.STATUS_DTYPE)
            rc = self._lib.lzb_encode_batch(self._h, fmt, _C.byref(opt), blob.ctypes.data, in_off.ctypes.data, n,
                                            out.ctypes.data, out_off.ctypes.data, out_len.ctypes.data, st.ctypes.data)
            if rc != _native.RC_OK:
                raise RuntimeError(f"lzb_encode_batch failed rc={rc}: {self.last_error()}")
            short = st["code"] == _native.E_CAPACITY
            if not short.any():
                break
            caps = _np.where(short, st["a0"], caps).astype(_np.uint64)
        return [out[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes() for i in range(n)]

    def raw_new(self, fmt, lc, lp, pb, dict_size):
        """A decompress::raw decoder object whose DecoderState lives on this context's device."""
        return RawHandle(self, fmt, lc, lp, pb, dict_size)

    def decompress_one(self, fmt, data, options=None):
        """lzb_decompress_alloc: scan + decode (+ capacity retry for end-marker .lzma)."""
        opt = (options or decompress.Options())._native()
        buf = bytes(data)
        out = _C.c_void_p()
        out_len, consumed = _C.c_size_t(), _C.c_size_t()
        st = _native.Status()
        rc = self._lib.lzb_decompress_alloc(self._h, fmt, _C.byref(opt), buf, len(buf), _C.byref(out),
                                            _C.byref(out_len), _C.byref(consumed), _C.byref(st))
        if rc != _native.RC_OK:
            raise RuntimeError(f"lzb_decompress_alloc failed rc={rc}: {self.last_error()}")
        payload = _C.string_at(out, out_len.value) if out_len.value else b""
        self._lib.lzb_free(out)
        row = _np.zeros((), dtype=_native.STATUS_DTYPE)
        row["code"], row["kind"], row["a0"], row["a1"], row["a2"] = st.code, st.kind, st.a0, st.a1, st.a2
        disp = "" if st.code == 0 else _native.format_status(self._lib, st)
        return StreamResult(payload, consumed.value, row, disp)


class RawHandle:
    """One decompress::raw decoder object on the device (lzb_raw_create .. lzb_raw_destroy)."""

    def __init__(self, ctx, fmt, lc, lp, pb, dict_size):
        self._ctx, self._lib = ctx, ctx._lib
        h = _C.c_void_p()
        rc = self._lib.lzb_raw_create(ctx.handle, fmt, lc, lp, pb, dict_size, _C.byref(h))
        if rc != _native.RC_OK:
            raise error.InternalError(f"lzb_raw_create failed rc={rc}: {ctx.last_error()}")
        self._h = h

    def reset(self):
        if self._lib.lzb_raw_reset(self._h) != _native.RC_OK:
            raise RuntimeError(f"lzb_raw_reset failed: {self._ctx.last_error()}")

    def decompress(self, data, options=None):
        opt = (options or decompress.Options())._native()
        buf = bytes(data)
        out = _C.c_void_p()
        out_len, consumed = _C.c_size_t(), _C.c_size_t()
        st = _native.Status()
        rc = self._lib.lzb_raw_decompress(self._h, _C.byref(opt), buf, len(buf), _C.byref(out), _C.byref(out_len),
                                          _C.byref(consumed), _C.byref(st))
        if rc != _native.RC_OK:
            raise RuntimeError(f"lzb_raw_decompress failed rc={rc}: {self._ctx.last_error()}")
        payload = _C.string_at(out, out_len.value) if out_len.value else b""
        self._lib.lzb_free(out)
        row = _np.zeros((), dtype=_native.STATUS_DTYPE)
        row["code"], row["kind"], row["a0"], row["a1"], row["a2"] = st.code, st.kind, st.a0, st.a1, st.a2
        disp = "" if st.code == 0 else _native.format_status(self._lib, st)
        return StreamResult(payload, consumed.value, row, disp)

    def close(self):
        if getattr(self, "_h", None) and self._ctx.handle:
            self._lib.lzb_raw_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiContext(Context):
    """Several CUDA devices behind one handle (lzb_create_multi): `decode_batch` splits a host batch into contiguous
    stream ranges with equal compressed bytes, every device uploads, decodes and returns its own range
    (lzb_decode_batch_multi).  Everything else (single streams, encoders, scans) runs on the first device."""

    def __init__(self, devices=None):
        self._lib = _native.load()
        m = _C.c_void_p()
        if devices:
            arr = (_C.c_int * len(devices))(*devices)
            rc = self._lib.lzb_create_multi(_C.byref(m), arr, len(devices))
        else:
            rc = self._lib.lzb_create_multi(_C.byref(m), None, 0)
        if rc != _native.RC_OK:
            raise RuntimeError(f"lzb_create_multi failed (rc={rc}): no usable CUDA device -- lzma_rs_b200 has no CPU fallback")
        self._m = m
        self._h = _C.c_void_p(self._lib.lzb_multi_ctx(m, 0))
        self.device_count = self._lib.lzb_multi_device_count(m)
        self.last_split = None

    def close(self):
        if getattr(self, "_m", None):
            self._lib.lzb_destroy_multi(self._m)
            self._m = self._h = None

    def _decode_call(self, fmt, opt, blob, in_off, n, out, out_off, out_len, consumed, st):
        split = _np.zeros(self.device_count + 1, dtype=_np.uint32)
        rc = self._lib.lzb_decode_batch_multi(self._m, fmt, _C.byref(opt), blob.ctypes.data, in_off.ctypes.data, n,
                                              out.ctypes.data, out_off.ctypes.data, out_len.ctypes.data,
                                              consumed.ctypes.data, st.ctypes.data, split.ctypes.data)
        self.last_split = split
        if rc != _native.RC_OK:
            raise RuntimeError(f"lzb_decode_batch_multi failed rc={rc}: {self._lib.lzb_multi_last_error(self._m).decode()}")


_default_ctx = None


def set_devices(devices=None):
    """Spread the module-level `*_decompress_batch` calls over these CUDA devices (None / empty: every visible device).
    Call before the first decode."""
    global _default_ctx
    if _default_ctx is not None:
        raise RuntimeError("set_devices() must precede the first decode")
    _default_ctx = MultiContext(devices)
    return _default_ctx


def _ctx():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def _read_all(inp):
    if isinstance(inp, (bytes, bytearray, memoryview)):
        return bytes(inp), None
    return inp.read(), inp


def _deliver(r, given, prefix, reader, output):
    """Writes the result to `output`, leaves unread trailing bytes in `reader`, raises the reference's error.
    given = bytes handed to the decoder, of which the first `prefix` were synthesised by the caller."""
    if output is not None and r.data:
        output.write(r.data)  # on error this is the reference's partial output
    if reader is not None and r.ok and hasattr(reader, "seek") and r.consumed != given:
        reader.seek(max(r.consumed, prefix) - given, 1)  # like BufRead::consume
    r.raise_for_status()
    return r.data


def _one(fmt, inp, output, options):
    data, reader = _read_all(inp)
    return _deliver(_ctx().decompress_one(fmt, data, options), len(data), 0, reader, output)


def lzma_decompress(input, output=None):
    """lzma_rs::lzma_decompress (src/lib.rs:44-49)."""
    return _one(_native.FMT_LZMA, input, output, None)


def lzma_decompress_with_options(input, output=None, options=None):
    """lzma_rs::lzma_decompress_with_options (src/lib.rs:52-60).  `allow_incomplete` is an option of the stream API only
    (options.rs:15-19): the one-shot decoder ignores it, here as in the reference."""
    if options is not None and options.allow_incomplete:
        options = decompress.Options(options.unpacked_size, options.memlimit, False)
    return _one(_native.FMT_LZMA, input, output, options)


def lzma2_decompress(input, output=None):
    """lzma_rs::lzma2_decompress (src/lib.rs:83-88)."""
    return _one(_native.FMT_LZMA2, input, output, None)


def xz_decompress(input, output=None):
    """lzma_rs::xz_decompress (src/lib.rs:100-105)."""
    return _one(_native.FMT_XZ, input, output, None)


def _compress(fmt, inp, output, options):
    data, _ = _read_all(inp)
    enc = _ctx().encode_batch(fmt, [data], options)[0]
    if output is not None:
        output.write(enc)
    return enc


def lzma_compress(input, output=None):
    """lzma_rs::lzma_compress (src/lib.rs:63-69): literal-only .lzma, like the reference's encoder."""
    return _compress(_native.FMT_LZMA, input, output, None)


def lzma_compress_with_options(input, output=None, options=None):
    """lzma_rs::lzma_compress_with_options (src/lib.rs:72-80)."""
    return _compress(_native.FMT_LZMA, input, output, options)


def lzma2_compress(input, output=None):
    """lzma_rs::lzma2_compress (src/lib.rs:91-97): stored chunks."""
    return _compress(_native.FMT_LZMA2, input, output, None)


def xz_compress(input, output=None):
    """lzma_rs::xz_compress (src/lib.rs:108-110): one block of stored chunks, no check."""
    return _compress(_native.FMT_XZ, input, output, None)


def lzma_compress_batch(datas, options=None):
    return _ctx().encode_batch(_native.FMT_LZMA, datas, options)


def lzma2_compress_batch(datas):
    return _ctx().encode_batch(_native.FMT_LZMA2, datas)


def xz_compress_batch(datas):
    return _ctx().encode_batch(_native.FMT_XZ, datas)


def lzma_decompress_batch(streams, options=None):
    return _ctx().decode_batch(_native.FMT_LZMA, streams, options)


def lzma2_decompress_batch(streams):
    return _ctx().decode_batch(_native.FMT_LZMA2, streams)


def xz_decompress_batch(streams):
    return _ctx().decode_batch(_native.FMT_XZ, streams)
