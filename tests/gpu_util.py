"""Helpers for the GPU tests: the product library driven through its C ABI (host and device entry points)."""
import ctypes as C

import numpy as np

from lzma_rs_b200 import Context, _native, decompress


def options_from(opts):
    mode = opts.get("unpacked_mode", 0)
    us = decompress.UnpackedSize(mode, opts.get("provided"))
    return decompress.Options(unpacked_size=us, memlimit=opts.get("memlimit"))


def host_decode(ctx, fmt, streams, opts):
    """lzb_decode_batch (host buffers), with the LZB_E_CAPACITY retry of Context.decode_batch."""
    return ctx.decode_batch(fmt, streams, options_from(opts))


class DeviceBatch:
    """Raw streams resident in device memory (torch tensors), decoded through lzb_batch_* / lzb_decode_batch_device."""

    def __init__(self, ctx, fmt, streams, capacities, opts=None):
        import torch
        self.ctx, self.fmt, self.n = ctx, fmt, len(streams)
        self.lib = _native.load()
        blob, self.in_off = _native.pack_streams(streams)
        caps = (np.asarray(capacities, dtype=np.uint64) + np.uint64(15)) // np.uint64(16) * np.uint64(16)
        self.out_off = np.zeros(self.n + 1, dtype=np.uint64)
        np.cumsum(caps, out=self.out_off[1:])
        self.d_in = torch.from_numpy(blob).cuda()
        self.d_out = torch.zeros(int(self.out_off[-1]) + 16, dtype=torch.uint8, device="cuda")
        self.opt = options_from(opts or {})._native()
        self.out_len = np.zeros(self.n, dtype=np.uint64)
        self.consumed = np.zeros(self.n, dtype=np.uint64)
        self.st = np.zeros(self.n, dtype=_native.STATUS_DTYPE)

    def decode(self):
        import torch
        rc = self.lib.lzb_decode_batch_device(self.ctx.handle, self.fmt, C.byref(self.opt), self.d_in.data_ptr(),
                                              self.in_off.ctypes.data, self.n, self.d_out.data_ptr(),
                                              self.out_off.ctypes.data, self.out_len.ctypes.data,
                                              self.consumed.ctypes.data, self.st.ctypes.data, None)
        assert rc == 0, (rc, self.ctx.last_error())
        torch.cuda.synchronize()
        return self

    def output(self, i):
        o = int(self.out_off[i])
        return self.d_out[o:o + int(self.out_len[i])].cpu().numpy().tobytes()

    def display(self, i):
        return "" if self.st[i]["code"] == 0 else _native.format_status(self.lib, self.st[i])


def host_decode_pinned(ctx, fmt, streams, capacities, opts=None, pin_input=True, blob_shift=0):
    """lzb_decode_batch with PINNED host buffers (torch.pin_memory): K1's mirror variant streams finished output pages
    to the host buffer itself, no device-to-host copy after the kernel.  Returns (list of bytes, out_len, consumed, st)."""
    import torch
    lib = _native.load()
    blob, in_off = _native.pack_streams(streams)
    if blob_shift:  # the batch starts at an odd offset of the caller's buffer (in_off[0] != 0)
        blob = np.concatenate([np.zeros(blob_shift, dtype=np.uint8), blob])
        in_off = in_off + np.uint64(blob_shift)
    n = len(streams)
    caps = (np.asarray(capacities, dtype=np.uint64) + np.uint64(15)) // np.uint64(16) * np.uint64(16)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(caps, out=out_off[1:])
    h_in = torch.from_numpy(blob).pin_memory() if pin_input else torch.from_numpy(blob)
    h_out = torch.full((int(out_off[-1]) + 16,), 0xEE, dtype=torch.uint8).pin_memory()
    out_len = np.zeros(n, dtype=np.uint64)
    consumed = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=_native.STATUS_DTYPE)
    opt = options_from(opts or {})._native()
    rc = lib.lzb_decode_batch(ctx.handle, fmt, C.byref(opt), h_in.data_ptr(), in_off.ctypes.data, n, h_out.data_ptr(),
                              out_off.ctypes.data, out_len.ctypes.data, consumed.ctypes.data, st.ctypes.data)
    assert rc == 0, (rc, ctx.last_error())
    hv = h_out.numpy()
    outs = [hv[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes() for i in range(n)]
    return outs, out_len, consumed, st
