"""GPU tier, C ABI v2: device-side sizing/layout, the peer decode (scatter + decode + gather in one
launch; exercised here on one GPU with same-device and pinned-host "remote" blobs), the one-process multi-device entry
point (two contexts on the same GPU), and the second-pass redo of the device path.  All against the oracle."""
import ctypes as C

import numpy as np
import pytest

import corpus
import oracle_py as oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lzma_rs_b200 import Context
    c = Context()
    yield c
    c.close()


def _mixed_batch(n, seed, lo=2_000, hi=150_000):
    rng = np.random.default_rng(seed)
    base_plain = [corpus.mixed_text(seed * 1000 + i, int(rng.integers(lo, hi))) for i in range(min(n, 64))]
    base = [corpus.raw_lzma2(p, dict_size=1 << 20) for p in base_plain]
    return [base[i % len(base)] for i in range(n)], [base_plain[i % len(base)] for i in range(n)]


def underconsumed_lclp4_stream():
    """An LZMA2 stream whose first chunk declares more packed bytes than the range decoder needs; the reference does not
    skip them (lzma2.rs:189-192) but parses them as the next chunk -- here a chunk with lc = 4, which the framing scan
    (skipping `packed` bytes) never sees."""
    e1 = corpus.LzmaEncoder(3, 0, 2)
    for b in b"first chunk, lc = 3":
        e1.literal(b)
    n1 = len(e1.hist)
    e2 = corpus.LzmaEncoder(4, 0, 2)
    for b in b"hidden chunk with lc = 4 behind the range coder's last byte":
        e2.literal(b)
    e2.match(9, 5)
    n2 = len(e2.hist)
    hidden = corpus.lzma2_chunk(e2.finish(), n2, 0xE0, corpus.props_byte(4, 0, 2))
    return corpus.lzma2_chunk(e1.finish() + hidden + b"\0", n1, 0xE0, corpus.props_byte(3, 0, 2)) + b"\0"


def test_scan_device_sizes_and_layout(ctx):
    """A device-resident batch is sized, laid out and decoded without its sizes ever visiting the host (only the total,
    to allocate the output blob): lzb_scan_device -> lzb_batch_prepare_device -> launch -> collect."""
    import torch
    from lzma_rs_b200 import _native
    lib = _native.load()
    streams, plains = _mixed_batch(300, 21)
    streams += [b"", b"\x00", streams[0][:50], corpus.stored_lzma2(b"stored bytes " * 999)]
    plains += [None] * 3 + [b"stored bytes " * 999]
    n = len(streams)
    blob, in_off = _native.pack_streams(streams)
    d_in = torch.from_numpy(blob).cuda()
    d_in_off = torch.from_numpy(in_off.astype(np.int64)).cuda()
    d_cap = torch.zeros(n, dtype=torch.int64, device="cuda")
    d_out_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    total = C.c_uint64(0)
    opt = _native.make_options()
    rc = lib.lzb_scan_device(ctx.handle, 1, C.byref(opt), d_in.data_ptr(), d_in_off.data_ptr(), n, d_cap.data_ptr(),
                             d_out_off.data_ptr(), C.byref(total), None)
    assert rc == 0, ctx.last_error()
    # same capacities as the host-side lzb_scan, and the layout is their 16-byte aligned exclusive prefix sum
    host_cap = np.zeros(n, dtype=np.uint64)
    assert lib.lzb_scan(ctx.handle, 1, C.byref(opt), blob.ctypes.data, in_off.ctypes.data, n, host_cap.ctypes.data) == 0
    cap = d_cap.cpu().numpy().astype(np.uint64)
    assert (cap == host_cap).all()
    off = d_out_off.cpu().numpy().astype(np.uint64)
    want_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((cap + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=want_off[1:])
    assert (off == want_off).all() and total.value == int(want_off[-1])
    d_out = torch.zeros(total.value + 16, dtype=torch.uint8, device="cuda")
    batch = C.c_void_p()
    rc = lib.lzb_batch_prepare_device(ctx.handle, 1, C.byref(opt), d_in.data_ptr(), d_in_off.data_ptr(), n, d_out.data_ptr(),
                                      d_out_off.data_ptr(), C.byref(batch))
    assert rc == 0, ctx.last_error()
    assert lib.lzb_batch_launch(batch, None) == 0
    out_len, cons = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=_native.STATUS_DTYPE)
    assert lib.lzb_batch_collect(batch, None, out_len.ctypes.data, cons.ctypes.data, st.ctypes.data) == 0
    lib.lzb_batch_destroy(batch)
    host = d_out.cpu().numpy()
    for i in range(n):
        ref = oracle.lzma2_decompress(streams[i])
        got = host[int(off[i]):int(off[i]) + int(out_len[i])].tobytes()
        assert got == ref.out, i
        disp = "" if st[i]["code"] == 0 else _native.format_status(lib, st[i])
        assert disp == ref.display, (i, disp, ref.display)
        if plains[i] is not None:
            assert got == plains[i]
    # .lzma headers on the device: known size, end-marker (heuristic capacity), header errors
    lz = [corpus.lzma_alone_known_size(plains[0]), corpus.lzma_alone(plains[1]), b"\x5d\x00", bytes([225]) + b"\0" * 12]
    blob, in_off = _native.pack_streams(lz)
    d_in = torch.from_numpy(blob).cuda()
    d_in_off = torch.from_numpy(in_off.astype(np.int64)).cuda()
    d_cap = torch.zeros(len(lz), dtype=torch.int64, device="cuda")
    rc = lib.lzb_scan_device(ctx.handle, 0, C.byref(opt), d_in.data_ptr(), d_in_off.data_ptr(), len(lz), d_cap.data_ptr(),
                             None, None, None)
    assert rc == 0
    host_cap = np.zeros(len(lz), dtype=np.uint64)
    assert lib.lzb_scan(ctx.handle, 0, C.byref(opt), blob.ctypes.data, in_off.ctypes.data, len(lz), host_cap.ctypes.data) == 0
    torch.cuda.synchronize()
    assert (d_cap.cpu().numpy().astype(np.uint64) == host_cap).all()


@pytest.mark.parametrize("where", ["device", "pinned_host"])
def test_peer_decode_one_launch(ctx, where):
    """lzb_decode_batch_peer: compressed bytes pulled from a blob that is NOT the context's staging memory (another
    device allocation / pinned host memory) behind the input gate, output pages stored to the destination blob by K1.
    A stream range that starts in the middle of both blobs, error streams mixed in, > 16 MiB so that the gate is armed."""
    import torch
    from lzma_rs_b200 import _native
    lib = _native.load()
    streams, plains = _mixed_batch(700, 33, 20_000, 120_000)
    streams[5] = streams[5][:len(streams[5]) // 2]
    streams[333] = b"\x55" + streams[333][1:]
    streams[400] = underconsumed_lclp4_stream()
    n = len(streams)
    assert sum(len(s) for s in streams) > 20 << 20
    refs = [oracle.lzma2_decompress(s) for s in streams]
    blob, in_off = _native.pack_streams(streams)
    caps = np.array([max(len(p), len(r.out)) for p, r in zip(plains, refs)], dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((caps + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
    if where == "device":
        src = torch.from_numpy(blob).cuda()
        dst = torch.full((int(out_off[-1]) + 16,), 0xEE, dtype=torch.uint8, device="cuda")
    else:
        src = torch.from_numpy(blob).pin_memory()
        dst = torch.full((int(out_off[-1]) + 16,), 0xEE, dtype=torch.uint8).pin_memory()
    opt = _native.make_options()
    lo, hi = 37, n - 11  # a rank's range: offsets are relative to the WHOLE blobs
    m = hi - lo
    out_len, cons = np.zeros(m, dtype=np.uint64), np.zeros(m, dtype=np.uint64)
    st = np.zeros(m, dtype=_native.STATUS_DTYPE)
    sub_in, sub_out = np.ascontiguousarray(in_off[lo:hi + 1]), np.ascontiguousarray(out_off[lo:hi + 1])
    rc = lib.lzb_decode_batch_peer(ctx.handle, 1, C.byref(opt), src.data_ptr(), sub_in.ctypes.data, m, dst.data_ptr(),
                                   sub_out.ctypes.data, out_len.ctypes.data, cons.ctypes.data, st.ctypes.data)
    assert rc == 0, ctx.last_error()
    torch.cuda.synchronize()
    host = dst.cpu().numpy()
    assert (host[:int(out_off[lo])] == 0xEE).all() and (host[int(out_off[hi]):] == 0xEE).all(), "wrote outside its range"
    for k in range(m):
        i = lo + k
        got = host[int(out_off[i]):int(out_off[i]) + int(out_len[k])].tobytes()
        disp = "" if st[k]["code"] == 0 else _native.format_status(lib, st[k])
        assert disp == refs[i].display, (i, disp, refs[i].display)
        assert got == refs[i].out, i
        if refs[i].ok:
            assert int(cons[k]) == refs[i].consumed


def test_multi_device_entry_point(ctx):
    """lzb_create_multi / lzb_decode_batch_multi: the host batch is split into contiguous stream ranges with equal
    compressed bytes, one thread and one context per listed device (here the same GPU twice and three times)."""
    import torch
    from lzma_rs_b200 import _native
    lib = _native.load()
    streams, plains = _mixed_batch(500, 44, 5_000, 90_000)
    streams[77] = streams[77][:100]
    n = len(streams)
    refs = [oracle.lzma2_decompress(s) for s in streams]
    blob, in_off = _native.pack_streams(streams)
    caps = np.array([len(p) for p in plains], dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((caps + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
    h_in = torch.from_numpy(blob).pin_memory()
    opt = _native.make_options()
    for devs in ([0, 0], [0, 0, 0], None):
        m = C.c_void_p()
        if devs is None:
            rc = lib.lzb_create_multi(C.byref(m), None, 0)
        else:
            arr = (C.c_int * len(devs))(*devs)
            rc = lib.lzb_create_multi(C.byref(m), arr, len(devs))
        assert rc == 0
        nd = lib.lzb_multi_device_count(m)
        assert nd == (len(devs) if devs else torch.cuda.device_count())
        h_out = torch.full((int(out_off[-1]) + 16,), 0xEE, dtype=torch.uint8).pin_memory()
        out_len, cons = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        st = np.zeros(n, dtype=_native.STATUS_DTYPE)
        split = np.zeros(nd + 1, dtype=np.uint32)
        rc = lib.lzb_decode_batch_multi(m, 1, C.byref(opt), h_in.data_ptr(), in_off.ctypes.data, n, h_out.data_ptr(),
                                        out_off.ctypes.data, out_len.ctypes.data, cons.ctypes.data, st.ctypes.data,
                                        split.ctypes.data)
        assert rc == 0, lib.lzb_multi_last_error(m)
        assert split[0] == 0 and split[-1] == n and (np.diff(split.astype(np.int64)) >= 0).all()
        if nd > 1:  # equal shares of the compressed bytes, to within one stream
            share = np.diff(in_off[split.astype(np.int64)].astype(np.int64))
            assert share.max() - share.min() <= 2 * max(len(s) for s in streams)
        hv = h_out.numpy()
        for i in range(n):
            got = hv[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes()
            disp = "" if st[i]["code"] == 0 else _native.format_status(lib, st[i])
            assert disp == refs[i].display and got == refs[i].out, i
        lib.lzb_destroy_multi(m)


def test_device_path_second_pass(ctx):
    """a stream whose later chunk (hidden from the framing scan behind an under-consumed chunk) needs larger
    lc+lp than the scan saw is decoded like the reference on the DEVICE entry points too (second pass at lc+lp = 4)."""
    import gpu_util
    s = underconsumed_lclp4_stream()
    ref = oracle.lzma2_decompress(s)
    assert ref.ok and b"hidden chunk" in ref.out
    filler, fp = _mixed_batch(40, 55)
    streams = filler[:20] + [s] + filler[20:]
    caps = [len(p) for p in fp[:20]] + [len(ref.out)] + [len(p) for p in fp[20:]]
    b = gpu_util.DeviceBatch(ctx, 1, streams, caps).decode()
    assert b.st[20]["code"] == 0, b.display(20)
    assert b.output(20) == ref.out and int(b.consumed[20]) == ref.consumed
    for i in (0, 19, 21, 40):
        assert b.st[i]["code"] == 0 and b.output(i) == (fp[:20] + [ref.out] + fp[20:])[i]
    r = gpu_util.host_decode(ctx, 1, [s], {})[0]
    assert r.ok and r.data == ref.out


def test_cpp_multi_client(ctx, tmp_path):
    """tests/cpp/multi_check.cpp: a compiled C++ client of lzb_create_multi / lzb_decode_batch_multi (pageable buffers)."""
    import os
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "multi_check"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "multi_check.cpp"), "-L" + os.path.join(root, "lzma_rs_b200"),
                           "-llzma_b200", "-Wl,-rpath," + os.path.join(root, "lzma_rs_b200"), "-o", str(exe)])
    plain = corpus.mixed_text(4242, 180_000)
    (tmp_path / "s.lzma2").write_bytes(corpus.raw_lzma2(plain, dict_size=1 << 20))
    (tmp_path / "s.plain").write_bytes(plain)
    for devs in ["0,0", "all", "0,0,0,0"]:
        r = subprocess.run([str(exe), devs, str(tmp_path / "s.lzma2"), str(tmp_path / "s.plain"), "301"],
                           capture_output=True, text=True)
        assert r.returncode == 0, (devs, r.stdout, r.stderr)
        nd = {"0,0": 2, "0,0,0,0": 4}.get(devs, torch.cuda.device_count())
        assert r.stdout.startswith(f"devices {nd} split 0 "), r.stdout


def test_drain_mode_many_streams(ctx):
    """More than two rounds of streams with a pinned output: the plain kernels publish a done word per stream and the copy
    engine moves finished streams while later rounds decode (host API and peer form).  Error streams, stored-chunk-only
    streams (copy kernel, no done words), an unaligned output layout and the second-pass stream are mixed in."""
    import gpu_util
    import torch
    from lzma_rs_b200 import _native
    lib = _native.load()
    rng = np.random.default_rng(77)
    base_plain = [corpus.mixed_text(9100 + i, int(rng.integers(1_000, 9_000))) for i in range(96)]
    base = [corpus.raw_lzma2(p, dict_size=1 << 16) for p in base_plain]
    n = 9100
    streams = [base[i % 96] for i in range(n)]
    plains = [base_plain[i % 96] for i in range(n)]
    streams[11] = streams[11][:40]
    streams[5000] = b"\x33" + streams[5000][1:]
    streams[9000] = underconsumed_lclp4_stream()
    for i in (100, 4000, 9099):
        plains[i] = bytes([i & 0xFF]) * 70_001
        streams[i] = corpus.stored_lzma2(plains[i])
    special = {11, 5000, 9000}
    refs = {i: oracle.lzma2_decompress(streams[i]) for i in special}
    caps = [len(refs[i].out) + 64 if i in special else len(plains[i]) for i in range(n)]
    # host API, pinned buffers, 16-byte aligned layout
    outs, out_len, consumed, st = gpu_util.host_decode_pinned(ctx, 1, streams, caps)
    for i in range(n):
        if i in special:
            disp = "" if st[i]["code"] == 0 else _native.format_status(lib, st[i])
            assert disp == refs[i].display and outs[i] == refs[i].out, i
        else:
            assert st[i]["code"] == 0 and outs[i] == plains[i] and int(consumed[i]) == len(streams[i]), i
    # peer form into a device blob whose stream regions are NOT 16-byte aligned (tight layout)
    blob, in_off = _native.pack_streams(streams)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(np.asarray(caps, dtype=np.uint64), out=out_off[1:])
    src = torch.from_numpy(blob).cuda()
    dst = torch.full((int(out_off[-1]) + 16,), 0xEE, dtype=torch.uint8, device="cuda")
    opt = _native.make_options()
    ol, cs = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
    st2 = np.zeros(n, dtype=_native.STATUS_DTYPE)
    rc = lib.lzb_decode_batch_peer(ctx.handle, 1, C.byref(opt), src.data_ptr(), in_off.ctypes.data, n, dst.data_ptr(),
                                   out_off.ctypes.data, ol.ctypes.data, cs.ctypes.data, st2.ctypes.data)
    assert rc == 0, ctx.last_error()
    torch.cuda.synchronize()
    host = dst.cpu().numpy()
    assert (st2["code"] == st["code"]).all() and (ol == out_len).all()
    for i in range(n):
        assert host[int(out_off[i]):int(out_off[i]) + int(ol[i])].tobytes() == outs[i], i


def test_python_multi_context(ctx):
    """lzma_rs_b200.MultiContext: the Python face of lzb_decode_batch_multi (same results as one context)."""
    import lzma_rs_b200 as L
    streams, plains = _mixed_batch(200, 66, 3_000, 60_000)
    streams.append(streams[0][:33])
    mc = L.MultiContext([0, 0, 0])
    try:
        got = mc.decode_batch(1, streams)
        want = ctx.decode_batch(1, streams)
        assert mc.device_count == 3 and mc.last_split[0] == 0 and mc.last_split[-1] == len(streams)
        for g, w, p in zip(got, want, plains + [None]):
            assert g.data == w.data and g.display == w.display and g.consumed == w.consumed
            if p is not None:
                assert g.ok and g.data == p
        assert not got[-1].ok and got[-1].display == "io error: failed to fill whole buffer"
    finally:
        mc.close()


def test_kernel_family_selection(ctx, monkeypatch):
    """Which K1 variant the planner picks (lzb_batch_kernel_name): batches of at most 8 streams per SM run on the latency
    kernels, LZB_NO_LAT=1 and larger batches on the throughput kernels; all of them bit-exact."""
    import ctypes as C

    import torch

    import corpus
    from lzma_rs_b200 import _native
    lib = _native.load()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    plain = [corpus.mixed_text(7100 + i, 3000 + 17 * i) for i in range(16)]
    comp = [corpus.raw_lzma2(p, dict_size=1 << 16) for p in plain]

    def run(n, env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        streams = [comp[i % 16] for i in range(n)]
        blob, in_off = _native.pack_streams(streams)
        sizes = np.array([len(plain[i % 16]) for i in range(n)], dtype=np.uint64)
        out_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum((sizes + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
        d_in = torch.from_numpy(blob).cuda()
        d_out = torch.zeros(int(out_off[-1]) + 16, dtype=torch.uint8, device="cuda")
        opt = _native.make_options()
        batch = C.c_void_p()
        rc = lib.lzb_batch_prepare(ctx.handle, _native.FMT_LZMA2, C.byref(opt), d_in.data_ptr(), in_off.ctypes.data, n,
                                   d_out.data_ptr(), out_off.ctypes.data, C.byref(batch))
        assert rc == 0, (rc, ctx.last_error())
        name = lib.lzb_batch_kernel_name(batch).decode()
        assert lib.lzb_batch_launch(batch, None) == 0
        st = np.zeros(n, dtype=_native.STATUS_DTYPE)
        ol = np.zeros(n, dtype=np.uint64)
        cs = np.zeros(n, dtype=np.uint64)
        assert lib.lzb_batch_collect(batch, None, ol.ctypes.data, cs.ctypes.data, st.ctypes.data) == 0
        lib.lzb_batch_destroy(batch)
        for k in env:
            monkeypatch.delenv(k)
        assert (st["code"] == 0).all() and (ol == sizes).all()
        out = d_out.cpu().numpy()
        for i in (0, 1, n // 2, n - 1):
            assert out[int(out_off[i]):int(out_off[i]) + int(ol[i])].tobytes() == plain[i % 16], (name, i)
        return name

    assert run(5, {}) == "lzb_decode_lat_kernel"
    assert run(8 * sms, {}) == "lzb_decode_lat_kernel"
    assert run(5, {"LZB_NO_LAT": "1"}) == "lzb_decode_kernel"
    assert run(8 * sms + 1, {}) == "lzb_decode_kernel"
    assert run(28 * sms + 300, {}) in ("lzb_decode_kernel", "lzb_decode_sched_kernel")
