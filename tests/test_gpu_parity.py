"""GPU tier: the CUDA path (through the C ABI) against the oracle, bit-exact, on the whole parity corpus, the
reference's golden vectors, and BASELINE-sized batches via size-independent properties."""
import hashlib
import io

import numpy as np
import pytest

import cases
import corpus
import oracle_py as oracle
import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from lzma_rs_b200 import Context
    c = Context()
    yield c
    c.close()


def _host(ctx):
    import gpu_util
    return lambda fmt, streams, opts: gpu_util.host_decode(ctx, fmt, streams, opts)


@pytest.fixture(params=["auto", "throughput"])
def kernel_family(request, monkeypatch):
    """Batches of up to 8 streams per SM run on the latency kernels (lzb_decode_lat_*); LZB_NO_LAT=1 sends them through
    the throughput kernels, so that the small batches of the parity corpus cover both families."""
    if request.param == "throughput":
        monkeypatch.setenv("LZB_NO_LAT", "1")
    return request.param


@pytest.mark.parametrize("family", ["valid_lzma2_cases", "valid_lzma_cases", "hand_encoded_cases",
                                    "truncation_and_corruption_cases", "xz_cases", "xz_chain_cases"])
def test_cuda_path_matches_oracle(ctx, family, kernel_family):
    bad, n = [], 0
    for (fmt, okey), named in parity.group_cases(getattr(cases, family)()).items():
        bad += parity.check_group(_host(ctx), fmt, dict(okey), named)
        n += len(named)
    assert not bad, f"{len(bad)}/{n} mismatches:\n" + "\n".join(bad[:40])


def test_golden_vectors_on_gpu(ctx, golden):
    """Every golden vector of the reference's own tests (tests/golden/, see make_golden.py) through the CUDA path."""
    fmt = {"lzma": 0, "xz": 2}
    for f in ("lzma", "xz"):
        vs = list(golden.vectors(fmt=f))
        res = _host(ctx)(fmt[f], [golden.compressed(v) for v in vs], {})
        for v, r in zip(vs, res):
            assert hashlib.sha256(r.data).hexdigest() == v["plain_sha256"], v["name"]
            if "error" in v:
                how, want = v["error_match"], v["error"]
                assert (r.display == want if how == "exact" else r.display.startswith(want) if how == "prefix"
                        else want in r.display), (v["name"], r.display)
            else:
                assert r.ok and r.consumed == v["length"], (v["name"], r.display)


def test_reference_api_mirror(ctx, golden):
    """lzma_rs::{lzma_decompress, lzma2_decompress, xz_decompress} mirrors: reader/writer roles and error types."""
    import lzma_rs_b200 as L
    v = next(x for x in golden.vectors() if x["name"] == "foo.txt.lzma")
    out = io.BytesIO()
    L.lzma_decompress(io.BytesIO(golden.compressed(v)), out)
    assert hashlib.sha256(out.getvalue()).hexdigest() == v["plain_sha256"]
    v = next(x for x in golden.vectors() if x["name"] == "good-1-lzma2-3.xz")
    assert hashlib.sha256(L.xz_decompress(golden.compressed(v))).hexdigest() == v["plain_sha256"]
    data = corpus.mixed_text(77, 100_000)
    rd = io.BytesIO(corpus.raw_lzma2(data) + b"tail")
    assert L.lzma2_decompress(rd) == data
    assert rd.read() == b"tail"  # only the consumed bytes were taken from the reader
    with pytest.raises(L.error.HeaderTooShort):
        L.lzma_decompress(b"")
    with pytest.raises(L.error.LzmaError, match="must be < 225"):
        L.lzma_decompress(b"\xff" * 32)
    v = next(x for x in golden.vectors() if x["name"] == "corrupt-footer.xz")
    out = io.BytesIO()
    with pytest.raises(L.error.XzError) as ei:
        L.xz_decompress(golden.compressed(v), out)
    assert str(ei.value) == "xz error: Invalid footer CRC32: expected 0x01234567 but got 0x8b0d303e"
    assert hashlib.sha256(out.getvalue()).hexdigest() == v["plain_sha256"]  # the valid block was already written
    opts = L.decompress.Options(unpacked_size=L.decompress.UnpackedSize.ReadHeaderButUseProvided(None), memlimit=0)
    with pytest.raises(L.error.LzmaError, match="exceeded memory limit of 0"):
        L.lzma_decompress_with_options(corpus.dumb_lzma(b"Some data"), None, opts)


def test_device_resident_batch(ctx):
    """lzb_decode_batch_device: inputs and outputs stay in HBM (what bench.py times)."""
    import gpu_util
    plains = [corpus.mixed_text(500 + i, int(s)) for i, s in enumerate([0, 1, 100, 65536, 65536, 200_000, 300_000, 4097])]
    comp = [corpus.raw_lzma2(p, dict_size=1 << 20) for p in plains]
    b = gpu_util.DeviceBatch(ctx, 1, comp, [len(p) for p in plains]).decode()
    for i, p in enumerate(plains):
        assert b.st[i]["code"] == 0, b.display(i)
        assert b.output(i) == p and int(b.consumed[i]) == len(comp[i])
    # .lzma on the device path + exact capacity too small -> LZB_E_CAPACITY
    comp1 = [corpus.lzma_alone_known_size(p, dict_size=1 << 16) for p in plains[2:6]]
    b = gpu_util.DeviceBatch(ctx, 0, comp1, [len(p) + 288 for p in plains[2:6]]).decode()
    for i, p in enumerate(plains[2:6]):
        assert b.st[i]["code"] == 0 and b.output(i) == p, b.display(i)
    b = gpu_util.DeviceBatch(ctx, 1, comp[3:5], [1000, 65536]).decode()
    assert b.st[0]["code"] == -1 and b.st[1]["code"] == 0


def test_device_crc(ctx):
    import ctypes as C
    import torch
    import zlib
    from lzma_rs_b200 import _native
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, size=3_000_000, dtype=np.uint8)
    d = torch.from_numpy(data).cuda()
    off = np.array([0, 1, 17, 4096, 5000, 100_003, 1_000_000, 0], dtype=np.uint64)
    ln = np.array([0, 1, 4095, 4096, 4097, 300_001, 2_000_000, 3_000_000], dtype=np.uint64)
    c32 = np.zeros(len(off), dtype=np.uint32)
    c64 = np.zeros(len(off), dtype=np.uint64)
    rc = _native.load().lzb_crc_device(ctx.handle, d.data_ptr(), off.ctypes.data, ln.ctypes.data, len(off),
                                       c32.ctypes.data, c64.ctypes.data, None)
    assert rc == 0
    for i in range(len(off)):
        seg = data[int(off[i]):int(off[i] + ln[i])].tobytes()
        assert int(c32[i]) == zlib.crc32(seg), i
        assert int(c64[i]) == oracle.crc64(seg), i


def test_config2_scale_properties(ctx):
    """BASELINE config 2 shape (many 64 KiB LZMA2 streams): round trip against the plaintexts the streams were made
    from + checksum-of-checksums against the oracle; 1024 streams keeps the oracle side to seconds."""
    import gpu_util
    n = 1024
    comp, plain = corpus.build_lzma2_corpus(2, n, lambda i: 65536, 1 << 18, distinct=256, threads=8)
    b = gpu_util.DeviceBatch(ctx, 1, comp, [len(p) for p in plain]).decode()
    assert (b.st["code"] == 0).all()
    h_gpu, h_ref = hashlib.sha256(), hashlib.sha256()
    for i in range(n):
        out = b.output(i)
        assert out == plain[i], i
        h_gpu.update(hashlib.sha256(out).digest())
    for i in range(0, n, 16):  # oracle on a sample (every 16th stream) + plaintext identity on all
        r = oracle.lzma2_decompress(comp[i])
        assert r.ok and r.out == plain[i]
    for i in range(n):
        h_ref.update(hashlib.sha256(plain[i]).digest())
    assert h_gpu.digest() == h_ref.digest()


def test_config5_rep0_stress(ctx):
    """BASELINE config 5: all-overlapping rep0 matches (dist=1, len=273)."""
    import gpu_util
    n = 256
    streams = [corpus.rep0_stress_lzma2(262144, byte=i & 0xFF) for i in range(n)]
    b = gpu_util.DeviceBatch(ctx, 1, streams, [262144] * n).decode()
    assert (b.st["code"] == 0).all()
    for i in range(n):
        assert b.output(i) == bytes([i & 0xFF]) * 262144


def test_config4_xz_multiblock(ctx):
    """BASELINE config 4 shape: multi-block .xz files with CRC32 block checks."""
    files, plains = [], []
    for i in range(32):
        p = corpus.mixed_text(4_000_000 + i, 1 << 20)
        plains.append(p)
        files.append(corpus.xz_file(p, block_size=1 << 18, check=corpus.CHECK_CRC32))
    res = _host(ctx)(2, files, {})
    for r, p, f in zip(res, plains, files):
        assert r.ok and r.data == p and r.consumed == len(f), r.display


def test_cpp_host_mirror(ctx, golden, tmp_path):
    """lzma_rs_b200/host/lzma_rs.hpp: the C++ mirror of the reference API, compiled with g++ against the C ABI."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host_check"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + root, "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "host_check.cpp"), "-L" + os.path.join(root, "lzma_rs_b200"),
                           "-llzma_b200", "-Wl,-rpath," + os.path.join(root, "lzma_rs_b200"), "-o", str(exe)])
    for name, fmt in [("foo.txt.lzma", "lzma"), ("good-1-lzma2-4.xz", "xz"), ("corrupt-footer.xz", "xz"),
                      ("foo.txt.lzma", "rawlzma"), ("range-coder-edge-case.lzma", "rawlzma")]:
        v = next(x for x in golden.vectors() if x["name"] == name)
        src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
        src.write_bytes(golden.compressed(v))
        r = subprocess.run([str(exe), fmt, str(src), str(dst)], capture_output=True, text=True)
        assert hashlib.sha256(dst.read_bytes()).hexdigest() == v["plain_sha256"], (name, fmt, r.returncode, r.stderr)
        if "error" in v:
            assert r.returncode == 3 and r.stderr == v["error"], (name, r.stderr)
        else:
            assert r.returncode == 0, (name, r.stderr)
    plain = corpus.mixed_text(31337, 150_000)
    src, dst = tmp_path / "b.lzma2", tmp_path / "b.out"  # batch form + set_devices of the C++ mirror
    src.write_bytes(corpus.raw_lzma2(plain, dict_size=1 << 20))
    r = subprocess.run([str(exe), "batch-lzma2", str(src), str(dst)], capture_output=True, text=True)
    assert r.returncode == 0 and dst.read_bytes() == plain, (r.returncode, r.stderr)
    for fmt in ("rt-lzma", "rt-lzma2", "rt-xz"):  # compress side of the C++ mirror: round trips
        src, dst = tmp_path / "plain.bin", tmp_path / "rt.bin"
        src.write_bytes(plain)
        r = subprocess.run([str(exe), fmt, str(src), str(dst)], capture_output=True, text=True)
        assert r.returncode == 0 and dst.read_bytes() == plain, (fmt, r.stderr)


def test_host_api_pinned_output_mirror(ctx):
    """Pinned host output: the mirror variant of K1 writes finished 4 KiB pages to the host while decoding."""
    import gpu_util
    sizes = [0, 1, 15, 16, 17, 4095, 4096, 4097, 8192, 65536, 100_001, 300_000, 1 << 20]
    plains = [corpus.mixed_text(900 + i, s) for i, s in enumerate(sizes)]
    comp = [corpus.raw_lzma2(p, dict_size=1 << 20) for p in plains]
    outs, out_len, consumed, st = gpu_util.host_decode_pinned(ctx, 1, comp, [len(p) for p in plains])
    assert (st["code"] == 0).all()
    for o, p, c, k in zip(outs, plains, comp, consumed):
        assert o == p and int(k) == len(c)
    # .lzma with errors (partial output = whole dict-size slabs) and .xz files through the same pinned path
    enc = corpus.LzmaEncoder(3, 0, 2)
    for i in range(10_000):
        enc.literal(i & 0xFF)
    enc.match(4, 9_000)  # beyond the 4 KiB dictionary
    enc.end_marker()
    bad = corpus.lzma_header(3, 0, 2, 4096) + enc.finish()
    ref = oracle.lzma_decompress(bad)
    outs, out_len, consumed, st = gpu_util.host_decode_pinned(ctx, 0, [bad, corpus.lzma_alone(plains[9])], [20_000, 70_000])
    assert outs[0] == ref.out and len(ref.out) == 8192 and st[0]["code"] != 0
    assert outs[1] == plains[9] and st[1]["code"] == 0
    files = [corpus.xz_file(plains[11], block_size=1 << 16, check=corpus.CHECK_CRC64), corpus.xz_file(plains[12], block_size=100_000)]
    outs, out_len, consumed, st = gpu_util.host_decode_pinned(ctx, 2, files, [len(plains[11]), len(plains[12])])
    assert (st["code"] == 0).all() and outs[0] == plains[11] and outs[1] == plains[12]


def test_fuzz_differential(ctx):
    """Differential fuzzing in the style of the reference's fuzz targets (fuzz/fuzz_targets/compare_xz.rs:28-37:
    both fail or both equal): thousands of mutated streams per format in ONE batch call each, against the oracle."""
    import random
    rnd = random.Random(20260925)
    seeds = {
        0: [corpus.lzma_alone(corpus.mixed_text(3000 + i, n), dict_size=d) for i, (n, d) in
            enumerate([(500, 4096), (6000, 4096), (20_000, 1 << 16)])] +
           [corpus.lzma_alone_known_size(corpus.mixed_text(3010, 3000), dict_size=4096)],
        1: [corpus.raw_lzma2(corpus.mixed_text(3100 + i, n), dict_size=d, lc=lc, lp=lp, pb=pb) for i, (n, d, lc, lp, pb) in
            enumerate([(700, 4096, 3, 0, 2), (9000, 1 << 16, 0, 2, 0), (150_000, 1 << 16, 4, 0, 4)])] +
           [corpus.stored_lzma2(corpus.mixed_text(3110, 2000))],
        2: [corpus.xz_file(corpus.mixed_text(3200, 5000), block_size=1500, check=corpus.CHECK_CRC32),
            corpus.xz_file(corpus.mixed_text(3201, 40_000), block_size=1 << 14, check=corpus.CHECK_CRC64, with_sizes=True)],
    }

    def mutate(b):
        b = bytearray(b)
        k = rnd.random()
        if k < 0.55:
            for _ in range(rnd.choice([1, 1, 1, 2, 4])):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        elif k < 0.7:
            b = b[:rnd.randrange(len(b))]
        elif k < 0.8:
            i = rnd.randrange(len(b))
            b[i:i] = bytes(rnd.randrange(256) for _ in range(rnd.choice([1, 2, 7])))
        elif k < 0.9:
            i = rnd.randrange(len(b))
            del b[i:i + rnd.choice([1, 2, 5])]
        else:
            i = rnd.randrange(min(len(b), 24))
            b[i] ^= 1 << rnd.randrange(8)
        return bytes(b)

    total = 0
    for fmt, srcs in seeds.items():
        named = [(f"fuzz-{fmt}-{i}", mutate(srcs[i % len(srcs)])) for i in range(1200)]
        bad = parity.check_group(_host(ctx), fmt, {}, named)
        assert not bad, f"fmt {fmt}: {len(bad)} mismatches:\n" + "\n".join(bad[:20])
        total += len(named)
    assert total == 3600


def test_batch_shape_extremes(ctx):
    """Many tiny streams next to a few long ones (scheduler imbalance), and a long .lzma whose 4 KiB ring wraps often."""
    import gpu_util
    tiny_plain = [corpus.mixed_text(7000 + i, 40 + (i % 200)) for i in range(512)]
    tiny = [corpus.raw_lzma2(p) for p in tiny_plain]
    big_plain = [corpus.mixed_text(7777, 6 << 20), b"\0" * (3 << 20) + corpus.mixed_text(7778, 1 << 20)]
    big = [corpus.raw_lzma2(p, dict_size=1 << 22) for p in big_plain]
    n_tiny = 20_000
    streams = [tiny[i % 512] for i in range(n_tiny)] + big
    plains = [tiny_plain[i % 512] for i in range(n_tiny)] + big_plain
    b = gpu_util.DeviceBatch(ctx, 1, streams, [len(p) for p in plains]).decode()
    assert (b.st["code"] == 0).all()
    for i in list(range(0, n_tiny, 997)) + [n_tiny, n_tiny + 1]:
        assert b.output(i) == plains[i], i
    assert int(b.out_len.sum()) == sum(len(p) for p in plains)
    wrap_plain = corpus.mixed_text(7900, 2 << 20)
    r = ctx.decode_batch(0, [corpus.lzma_alone(wrap_plain, dict_size=4096), corpus.lzma_alone_known_size(wrap_plain, dict_size=1 << 16)])
    assert all(x.ok and x.data == wrap_plain for x in r)


def test_size_mix_placement(ctx):
    """North-star shaped batch (sizes log-uniform over a 64x range, more streams than resident warps): the placement
    planner (lzb_sched.h) pre-assigns first items and parks warps next to the longest streams; every stream must still
    be decoded exactly once and bit-exact."""
    import emul_py
    import gpu_util
    rng = np.random.default_rng(99)
    base_plain = [corpus.mixed_text(8100 + i, int(2048 * 64 ** rng.random())) for i in range(600)]
    base = [corpus.raw_lzma2(p, dict_size=1 << 20) for p in base_plain]
    n = 6000
    streams = [base[i % 600] for i in range(n)]
    plains = [base_plain[i % 600] for i in range(n)]
    work = np.sort(np.array([len(s) for s in streams], dtype=np.float64))[::-1]
    _, info = emul_py.sched_plan(work)  # the same planner code, compiled for the CPU tier
    assert info["throttled"] and info["parked"] > 0, info
    b = gpu_util.DeviceBatch(ctx, 1, streams, [len(p) for p in plains]).decode()
    assert (b.st["code"] == 0).all()
    assert (b.consumed == np.array([len(s) for s in streams], dtype=np.uint64)).all()
    host = b.d_out.cpu().numpy()
    for i in range(n):
        o = int(b.out_off[i])
        assert host[o:o + int(b.out_len[i])].tobytes() == plains[i], i


def test_raw_decoders(ctx):
    """decompress::raw::{LzmaParams, LzmaDecoder, Lzma2Decoder} (feature raw_decoder, src/lib.rs:29-35) on the device:
    decoder objects whose DecoderState survives between decompress() calls (lzb_raw_*), against the oracle's decoder
    objects -- the same checks the CPU tier runs through the host emulation (tests/test_raw_header.py)."""
    from test_raw_header import raw_decoder_checks
    raw_decoder_checks(ctx)
    # a large-lc decoder (literal table of 0x300 << 8 entries in the state record) used three times in a row
    import lzma_rs_b200 as L
    raw = L.decompress.raw
    data = corpus.mixed_text(777, 6_000)
    enc = corpus.LzmaEncoder(8, 4, 0)  # liblzma refuses lc + lp > 4: hand-encoded payload (literals + a match every 40 bytes)
    i = 0
    while i < len(data):
        if i >= 64 and i % 40 == 0 and i + 12 <= len(data):
            k = data.rfind(data[i:i + 3], 0, i)
            if k >= 0 and data[k:k + 3] == data[i:i + 3]:
                n = 3
                while n < 12 and data[k + n] == data[i + n] and k + n < i:
                    n += 1
                enc.match(n, i - k)
                i += n
                continue
        enc.literal(data[i])
        i += 1
    assert bytes(enc.hist) == data
    payload = enc.finish()
    dec = raw.LzmaDecoder(raw.LzmaParams(raw.LzmaProperties(8, 4, 0), 1 << 16, len(data)), None, ctx)
    ora = oracle.RawDecoder(0, 8, 4, 0, 1 << 16, len(data))
    blob = b"\0" * 13 + payload
    assert ora.decompress(payload).out == data
    ora.reset()
    for _ in range(3):
        want = ora.decompress(blob[13:])
        try:
            got, disp = dec.decompress(blob[13:]), ""
        except L.error.Error as e:
            got, disp = None, str(e)
        assert disp == want.display and (got is None or got == want.out)
        if not want.ok:
            break


def test_host_api_gated_upload(ctx):
    """lzb_decode_batch with a pinned output and > 16 MiB of input: K1 is launched before the input blob has arrived
    and every warp waits for its own stream's bytes (input gate).  Pinned and pageable input, blob starting at an odd
    offset, streams that end in errors and preset (header) errors mixed in; compared with the ungated path."""
    import os
    import gpu_util
    rng = np.random.default_rng(11)
    base_plain = [corpus.mixed_text(8800 + i, int(rng.integers(20_000, 200_000))) for i in range(96)]
    base = [corpus.raw_lzma2(p, dict_size=1 << 20) for p in base_plain]
    n = 1200
    streams = [base[i % 96] for i in range(n)]
    plains = [base_plain[i % 96] for i in range(n)]
    streams[7] = streams[7][:len(streams[7]) // 2]        # truncated: UnexpectedEof
    streams[600] = b"\x55" + streams[600][1:]             # invalid LZMA2 status byte
    assert sum(len(s) for s in streams) > 40 << 20
    caps = [len(p) for p in plains]
    want = [oracle.lzma2_decompress(streams[i]) for i in (7, 600)]
    for pin_input, shift in ((True, 0), (True, 5), (False, 21)):
        outs, out_len, consumed, st = gpu_util.host_decode_pinned(ctx, 1, streams, caps, pin_input=pin_input, blob_shift=shift)
        for i in range(n):
            if i in (7, 600):
                w = want[(7, 600).index(i)]
                assert st[i]["code"] != 0 and outs[i] == w.out, i
            else:
                assert st[i]["code"] == 0 and outs[i] == plains[i] and int(consumed[i]) == len(streams[i]), i
    os.environ["LZB_NO_GATE"] = "1"  # same call, upload in stream order before the kernel
    try:
        outs2, _, _, st2 = gpu_util.host_decode_pinned(ctx, 1, streams, caps)
    finally:
        del os.environ["LZB_NO_GATE"]
    assert outs2 == outs and (st2["code"] == st["code"]).all()


def test_gpu_encoders(ctx):
    """Compress side: lzma_compress / lzma2_compress / xz_compress on the GPU write exactly the bytes
    of the oracle's restatement of the reference's encoders, for every option, at chunk-size edges and in ragged batches;
    and the reference's own round-trip tests (tests/lzma.rs:16-28,146-168, tests/lzma2.rs, tests/xz.rs:30-52) pass
    through GPU encode -> GPU decode."""
    import lzma
    import lzma_rs_b200 as L
    rng = np.random.default_rng(77)
    datas = [b"", b"a", b"Hello world", bytes(65535), bytes(65536), bytes(65537), b"\xff" * (1 << 20), bytes(1 << 20),
             corpus.mixed_text(901, 200_000), rng.bytes(3 * 65536 + 17)]
    datas += [corpus.mixed_text(1000 + i, int(rng.integers(0, 150_000))) for i in range(120)]
    CO, US = L.compress.Options, L.compress.UnpackedSize
    got = ctx.encode_batch(1, datas)
    assert got == [oracle.lzma2_compress(d) for d in datas]
    got = ctx.encode_batch(2, datas)
    assert got == [oracle.xz_compress(d) for d in datas]
    assert all(lzma.decompress(x, format=lzma.FORMAT_XZ) == d for x, d in zip(got[:12], datas[:12]))
    small = datas[:6] + datas[8:40]
    got = ctx.encode_batch(0, small)
    assert got == [oracle.lzma_compress(d) for d in small]
    got = ctx.encode_batch(0, small, CO(US.SkipWritingToHeader()))
    assert got == [oracle.lzma_compress(d, skip_size_field=True) for d in small]
    got = ctx.encode_batch(0, small[:8], CO(US.WriteToHeader(1234)))
    assert got == [oracle.lzma_compress(d, value=1234) for d in small[:8]]
    # capacity too small is a per-stream status, the neighbours are untouched
    from lzma_rs_b200 import _native
    import ctypes as C
    blob, in_off = _native.pack_streams([b"x" * 1000, b"y" * 1000, b"z" * 1000])
    out_off = np.array([0, 1008, 1008 + 500, 1008 + 500 + 1008], dtype=np.uint64)
    out = np.zeros(int(out_off[-1]) + 16, dtype=np.uint8)
    ol, st = np.zeros(3, dtype=np.uint64), np.zeros(3, dtype=_native.STATUS_DTYPE)
    assert _native.load().lzb_encode_batch(ctx.handle, 1, None, blob.ctypes.data, in_off.ctypes.data, 3, out.ctypes.data,
                                           out_off.ctypes.data, ol.ctypes.data, st.ctypes.data) == 0
    assert st["code"].tolist() == [0, _native.E_CAPACITY, 0] and int(st[1]["a0"]) == 1004 and ol.tolist() == [1004, 0, 1004]
    assert out[1008:1508].tobytes() == bytes(500)
    # round trips through the public API, the reference's own test inputs
    for d in (b"", b"Hello world", bytes(1 << 20), b"\xff" * (1 << 20), datas[8]):
        assert L.lzma2_decompress(L.lzma2_compress(d)) == d
        assert L.xz_decompress(L.xz_compress(d)) == d
    for d in (b"", b"Hello world", datas[8]):
        assert L.lzma_decompress(L.lzma_compress(d)) == d
        enc = L.lzma_compress_with_options(d, None, CO(US.WriteToHeader(len(d))))
        assert L.lzma_decompress(enc) == d
        enc = L.lzma_compress_with_options(d, None, CO(US.SkipWritingToHeader()))
        opts = L.decompress.Options(unpacked_size=L.decompress.UnpackedSize.UseProvided(len(d)))
        assert L.lzma_decompress_with_options(enc, None, opts) == d


def test_stored_only_route(ctx):
    """LZMA2 streams made of stored chunks only (what the reference's own encoder writes) are decoded by the copy kernel
    (lzb_stored_decode_kernel) instead of K1; everything else in the same batch still goes through K1.  Device path."""
    import struct
    import gpu_util
    rng = np.random.default_rng(3)

    def stored(data, sizes, first_status=1, other_status=2):
        out, pos = bytearray(), 0
        for k, n in enumerate(sizes):
            out += bytes([first_status if k == 0 else other_status]) + struct.pack(">H", n - 1) + data[pos:pos + n]
            pos += n
        assert pos == len(data)
        return bytes(out) + b"\0"

    big = rng.bytes(5 * 65536 + 123)
    odd = rng.bytes(70_001)
    streams = [
        corpus.stored_lzma2(big),                                        # reference encoder shape
        stored(odd, [1, 65536, 2, 4461, 1]),                              # ragged chunk sizes, status 1 then 2
        stored(odd[:10], [10], first_status=2),                           # never a dict reset: allowed
        corpus.stored_lzma2(big)[:-1],                                    # terminator missing -> K1 reports the error
        corpus.stored_lzma2(odd) + b"trailing bytes",                     # consumed stops after the 0x00
        corpus.raw_lzma2(corpus.mixed_text(4, 100_000), dict_size=1 << 20),  # ordinary compressed stream
        b"\0",                                                            # empty stream
        corpus.stored_lzma2(big),                                         # capacity too small below
    ]
    caps = [len(big), len(odd), 10, len(big), len(odd), 100_000, 0, 1000]
    b = gpu_util.DeviceBatch(ctx, 1, streams, caps).decode()
    for i, s in enumerate(streams):
        w = oracle.lzma2_decompress(s)
        if i == 7:
            assert b.st[i]["code"] == -1  # LZB_E_CAPACITY from K1
            continue
        assert (b.st[i]["code"] == 0) == w.ok, (i, b.display(i), w.display)
        assert b.output(i) == w.out, i
        if w.ok:
            assert int(b.consumed[i]) == w.consumed, i
        else:
            assert b.display(i) == w.display, i
    # a large all-stored batch, every byte checked
    datas = [rng.bytes(int(rng.integers(1, 400_000))) for _ in range(64)]
    streams = [corpus.stored_lzma2(d) for d in datas] * 8
    b = gpu_util.DeviceBatch(ctx, 1, streams, [len(d) for d in datas] * 8).decode()
    assert (b.st["code"] == 0).all()
    for i in range(len(streams)):
        assert b.output(i) == datas[i % 64] and int(b.consumed[i]) == len(streams[i])


def test_fuzz_regressions_on_gpu(ctx):
    """Inputs on which the fuzz soak once found a mismatch (e.g. an .xz block whose real end differs from the framing
    scan's prediction, followed by the index: the terminal had been read at the predicted offset)."""
    from test_emul_parity import fuzz_regressions
    for fmt, named in fuzz_regressions().items():
        bad = parity.check_group(_host(ctx), fmt, {}, named)
        assert not bad, "\n".join(bad)


def test_concurrent_batches_on_two_streams(ctx):
    """Two prepared device batches of one context in flight at the same time on different CUDA streams: every batch owns
    its result and workspace buffers (include/lzma_b200.h), so neither disturbs the other."""
    import ctypes as C
    import torch
    from lzma_rs_b200 import _native
    lib = _native.load()
    opt = _native.make_options()
    sets = []
    for k in range(2):
        plains = [corpus.mixed_text(9100 + 50 * k + i, 60_000 + 7000 * i) for i in range(40)]
        comp = [corpus.raw_lzma2(p, dict_size=1 << 20) for p in plains] * 8
        plains = plains * 8
        blob, in_off = _native.pack_streams(comp)
        sizes = np.array([len(p) for p in plains], dtype=np.uint64)
        out_off = np.zeros(len(plains) + 1, dtype=np.uint64)
        np.cumsum((sizes + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=out_off[1:])
        d_in = torch.from_numpy(blob).cuda()
        d_out = torch.zeros(int(out_off[-1]) + 16, dtype=torch.uint8, device="cuda")
        batch = C.c_void_p()
        assert lib.lzb_batch_prepare(ctx.handle, 1, C.byref(opt), d_in.data_ptr(), in_off.ctypes.data, len(plains),
                                     d_out.data_ptr(), out_off.ctypes.data, C.byref(batch)) == 0
        sets.append((plains, in_off, out_off, d_in, d_out, batch, torch.cuda.Stream()))
    for _ in range(3):  # both batches are enqueued before either is waited for
        for s in sets:
            assert lib.lzb_batch_launch(s[5], C.c_void_p(s[6].cuda_stream)) == 0
    for plains, in_off, out_off, d_in, d_out, batch, stream in sets:
        n = len(plains)
        ol, cs, st = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=_native.STATUS_DTYPE)
        assert lib.lzb_batch_collect(batch, C.c_void_p(stream.cuda_stream), ol.ctypes.data, cs.ctypes.data, st.ctypes.data) == 0
        assert (st["code"] == 0).all()
        host = d_out.cpu().numpy()
        for i in range(n):
            o = int(out_off[i])
            assert host[o:o + int(ol[i])].tobytes() == plains[i], i
    for s in sets:
        lib.lzb_batch_destroy(s[5])


def test_reference_suite_on_gpu(ctx, golden):
    """tests/lzma.rs, tests/lzma2.rs, tests/xz.rs of the reference, restated over the public Python mirror, on the device."""
    import lzma_rs_b200 as L
    from test_reference_suite import run_reference_suite
    v = next(x for x in golden.vectors() if x["name"] == "foo.txt.lzma")
    foo = L.lzma_decompress(golden.compressed(v))
    assert hashlib.sha256(foo).hexdigest() == v["plain_sha256"]
    run_reference_suite(L, foo)


def test_structured_fuzz_on_gpu(ctx, kernel_family):
    """Structure-aware differential fuzz (tools/fuzz_soak.py generators, fixed seed) through the CUDA path."""
    from test_emul_parity import structured_fuzz
    bad = structured_fuzz(_host(ctx), 20261018, 1500)
    assert not bad, f"{len(bad)} mismatches:\n" + "\n".join(bad[:20])


def test_stream_facade(ctx):
    """decompress::Stream façade on the device: the reference's own stream tests (src/decode/stream.rs:348-499),
    including Options::allow_incomplete (known answer: half of small.txt's compressed bytes -> its first 26 bytes)."""
    from test_stream_facade import check_stream_facade
    check_stream_facade(ctx)
