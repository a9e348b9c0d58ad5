//! `lzma_b200`: the decode entry points of `lzma_rs` with the same signatures, executed on a B200 through the C ABI
//! of `include/lzma_b200.h`.  Replace `use lzma_rs::{lzma_decompress, lzma2_decompress, xz_decompress}` with
//! `use lzma_b200::{...}`; `decompress::{Options, UnpackedSize}` and `error::{Error, Result}` keep their shapes.
//!
//! Besides the one-shot functions there is a batch form of each (`*_decompress_batch(&[&[u8]])`, the call the benchmark
//! drives), a slice form that reports `consumed` (`decompress_slice`), and `set_devices` to spread batches over several
//! GPUs of the node.
//!
//! NOT COMPILED in the build image (no Rust toolchain there) -- see INTEGRATION.md.
use std::io;
use std::os::raw::{c_char, c_int, c_void};
use std::sync::Mutex;

pub mod error {
    use std::{fmt, io, result};
    /// Same variants as `lzma_rs::error::Error` (src/error.rs:7-17).
    #[derive(Debug)]
    pub enum Error {
        IoError(io::Error),
        HeaderTooShort(io::Error),
        LzmaError(String),
        XzError(String),
    }
    pub type Result<T> = result::Result<T, Error>;
    impl From<io::Error> for Error {
        fn from(e: io::Error) -> Error {
            Error::IoError(e)
        }
    }
    impl fmt::Display for Error {
        fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
            match self {
                Error::IoError(e) => write!(f, "io error: {}", e),
                Error::HeaderTooShort(e) => write!(f, "header too short: {}", e),
                Error::LzmaError(e) => write!(f, "lzma error: {}", e),
                Error::XzError(e) => write!(f, "xz error: {}", e),
            }
        }
    }
    impl std::error::Error for Error {}
}

pub mod decompress {
    /// `lzma_rs::decompress::UnpackedSize` (src/decode/options.rs:24-43)
    #[derive(Clone, Copy, Debug, PartialEq, Eq, Default)]
    pub enum UnpackedSize {
        #[default]
        ReadFromHeader,
        ReadHeaderButUseProvided(Option<u64>),
        UseProvided(Option<u64>),
    }
    /// `lzma_rs::decompress::Options` (src/decode/options.rs:3-20)
    #[derive(Clone, Copy, Debug, PartialEq, Eq, Default)]
    pub struct Options {
        pub unpacked_size: UnpackedSize,
        pub memlimit: Option<usize>,
        pub allow_incomplete: bool,
    }
}

// ---- C ABI (include/lzma_b200.h) ------------------------------------------------------------------------------
#[repr(C)]
#[derive(Default, Clone, Copy)]
struct LzbOptions {
    unpacked_mode: u8,
    has_provided: u8,
    has_memlimit: u8,
    allow_incomplete: u8,
    reserved: [u8; 4],
    provided: u64,
    memlimit: u64,
}
#[repr(C)]
#[derive(Default, Clone, Copy)]
struct LzbStatus {
    code: i32,
    kind: i32,
    a0: u64,
    a1: u64,
    a2: u64,
}
#[repr(C)]
#[derive(Default, Clone, Copy)]
struct LzbCompressOptions {
    skip_size_field: u8,
    has_value: u8,
    reserved: [u8; 6],
    value: u64,
}
#[repr(C)]
struct LzbCtx {
    _private: [u8; 0],
}
#[repr(C)]
struct LzbMulti {
    _private: [u8; 0],
}
extern "C" {
    fn lzb_create(ctx: *mut *mut LzbCtx, device: c_int) -> c_int;
    fn lzb_create_multi(m: *mut *mut LzbMulti, dev_ids: *const c_int, n_dev: c_int) -> c_int;
    fn lzb_multi_ctx(m: *mut LzbMulti, k: c_int) -> *mut LzbCtx;
    fn lzb_scan(
        ctx: *mut LzbCtx, fmt: c_int, opt: *const LzbOptions, input: *const u8, in_off: *const u64, n: u32,
        capacity: *mut u64,
    ) -> c_int;
    fn lzb_decode_batch(
        ctx: *mut LzbCtx, fmt: c_int, opt: *const LzbOptions, input: *const u8, in_off: *const u64, n: u32, out: *mut u8,
        out_off: *const u64, out_len: *mut u64, consumed: *mut u64, st: *mut LzbStatus,
    ) -> c_int;
    fn lzb_decode_batch_multi(
        m: *mut LzbMulti, fmt: c_int, opt: *const LzbOptions, input: *const u8, in_off: *const u64, n: u32, out: *mut u8,
        out_off: *const u64, out_len: *mut u64, consumed: *mut u64, st: *mut LzbStatus, split: *mut u32,
    ) -> c_int;
    fn lzb_decompress_alloc(
        ctx: *mut LzbCtx, fmt: c_int, opt: *const LzbOptions, input: *const u8, in_len: usize, out: *mut *mut u8,
        out_len: *mut usize, consumed: *mut usize, st: *mut LzbStatus,
    ) -> c_int;
    fn lzb_free(p: *mut c_void);
    fn lzb_encode_bound(fmt: c_int, opt: *const LzbCompressOptions, in_len: u64) -> u64;
    fn lzb_encode_batch(
        ctx: *mut LzbCtx, fmt: c_int, opt: *const LzbCompressOptions, input: *const u8, in_off: *const u64, n: u32,
        out: *mut u8, out_off: *const u64, out_len: *mut u64, st: *mut LzbStatus,
    ) -> c_int;
    fn lzb_format_error(st: *const LzbStatus, buf: *mut c_char, buf_len: usize) -> usize;
}
const FMT_LZMA: c_int = 0;
const FMT_LZMA2: c_int = 1;
const FMT_XZ: c_int = 2;

/// The process-wide engine: one context on the current device, or (after `set_devices`) one per listed device.
struct Ctx {
    one: *mut LzbCtx,
    multi: *mut LzbMulti, // null unless set_devices() was called with more than one device
}
unsafe impl Send for Ctx {}
static CTX: Mutex<Option<Ctx>> = Mutex::new(None);

/// Use these CUDA devices for the `*_batch` functions (host batches are split into contiguous stream ranges with equal
/// compressed bytes, every device uploads and decodes its own range: `lzb_decode_batch_multi`).  Call before the first
/// decode; an empty slice = every visible device.  The single-stream functions use the first device.
pub fn set_devices(devices: &[i32]) -> io::Result<()> {
    let mut g = CTX.lock().unwrap();
    if g.is_some() {
        return Err(io::Error::new(io::ErrorKind::Other, "lzma_b200: set_devices() must precede the first decode"));
    }
    let mut m: *mut LzbMulti = std::ptr::null_mut();
    let ids: Vec<c_int> = devices.iter().map(|&d| d as c_int).collect();
    let rc = unsafe { lzb_create_multi(&mut m, if ids.is_empty() { std::ptr::null() } else { ids.as_ptr() }, ids.len() as c_int) };
    if rc != 0 {
        return Err(io::Error::new(io::ErrorKind::Other, format!("lzb_create_multi failed ({})", rc)));
    }
    *g = Some(Ctx { one: unsafe { lzb_multi_ctx(m, 0) }, multi: m });
    Ok(())
}

fn with_engine<T>(f: impl FnOnce(&Ctx) -> T) -> io::Result<T> {
    let mut g = CTX.lock().unwrap();
    if g.is_none() {
        let mut p: *mut LzbCtx = std::ptr::null_mut();
        let rc = unsafe { lzb_create(&mut p, -1) };
        if rc != 0 {
            return Err(io::Error::new(io::ErrorKind::Other, format!("lzb_create failed ({}): no CUDA device", rc)));
        }
        *g = Some(Ctx { one: p, multi: std::ptr::null_mut() });
    }
    Ok(f(g.as_ref().unwrap()))
}

fn with_ctx<T>(f: impl FnOnce(*mut LzbCtx) -> T) -> io::Result<T> {
    with_engine(|e| f(e.one))
}

fn to_error(st: &LzbStatus) -> error::Error {
    // message = the reference's Display string without its "xxx error: " prefix
    let mut buf = vec![0u8; 512];
    let n = unsafe { lzb_format_error(st, buf.as_mut_ptr() as *mut c_char, buf.len()) }.min(buf.len() - 1);
    let full = String::from_utf8_lossy(&buf[..n]).into_owned();
    let strip = |p: &str| full.strip_prefix(p).unwrap_or(&full).to_string();
    let eof = || io::Error::new(io::ErrorKind::UnexpectedEof, "failed to fill whole buffer");
    match st.kind {
        1 => error::Error::IoError(eof()),
        2 => error::Error::HeaderTooShort(eof()),
        3 => error::Error::LzmaError(strip("lzma error: ")),
        4 => error::Error::XzError(strip("xz error: ")),
        _ => error::Error::IoError(io::Error::new(io::ErrorKind::Other, full)),
    }
}

/// One decode of `data` (a complete stream, or a prefix of one).  Returns (status, output, consumed).
fn decode_slice(fmt: c_int, opt: &LzbOptions, data: &[u8]) -> error::Result<(LzbStatus, Vec<u8>, usize)> {
    let mut out: *mut u8 = std::ptr::null_mut();
    let (mut out_len, mut consumed) = (0usize, 0usize);
    let mut st = LzbStatus::default();
    let rc = with_ctx(|c| unsafe {
        lzb_decompress_alloc(c, fmt, opt, data.as_ptr(), data.len(), &mut out, &mut out_len, &mut consumed, &mut st)
    })?;
    if rc != 0 {
        return Err(error::Error::IoError(io::Error::new(io::ErrorKind::Other, format!("lzma_b200 call failed: {}", rc))));
    }
    let v = if out_len > 0 { unsafe { std::slice::from_raw_parts(out, out_len) }.to_vec() } else { Vec::new() };
    unsafe { lzb_free(out as *mut c_void) };
    Ok((st, v, consumed))
}

/// Slice form of the three entry points: decodes the stream at the start of `data` and returns the output together with
/// the number of input bytes the reference would have consumed (`lzma_rs` leaves the rest of its reader unread:
/// src/decode/lzma2.rs:59-68, src/decode/lzma.rs:442-445).
pub fn decompress_slice(fmt: Format, data: &[u8], opts: &decompress::Options) -> error::Result<(Vec<u8>, usize)> {
    let (st, out, consumed) = decode_slice(fmt as c_int, &options(opts), data)?;
    if st.code != 0 {
        return Err(to_error(&st));
    }
    Ok((out, consumed))
}
/// Stream formats = the reference's three entry points.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum Format {
    Lzma = 0,
    Lzma2 = 1,
    Xz = 2,
}

/// Status codes that mean "the decoder ran out of input" (include/lzma_b200.h): the stream may simply continue
/// behind what the reader's buffer showed us.
fn ran_out_of_input(st: &LzbStatus) -> bool {
    matches!(st.code, 1 | 2 | 4 | 12 | 14 | 15 | 16 | 19 | 20)
}

/// End of a raw LZMA2 stream inside `buf` (index behind its 0x00 control byte) if the chunk headers seen so far reach it;
/// `None` = more bytes are needed.  Framing only (lzma2.rs:59-78, 128-136, 204-207); malformed framing ends the walk
/// where the reference's error would.
fn lzma2_extent(buf: &[u8]) -> Option<usize> {
    let mut q = 0usize;
    loop {
        let status = *buf.get(q)?;
        q += 1;
        if status == 0 {
            return Some(q);
        }
        if status == 1 || status == 2 {
            let n = ((*buf.get(q)? as usize) << 8 | *buf.get(q + 1)? as usize) + 1;
            q += 2 + n;
        } else if status < 0x80 {
            return Some(q); // invalid control byte: the reference fails here
        } else {
            let packed = ((*buf.get(q + 2)? as usize) << 8 | *buf.get(q + 3)? as usize) + 1;
            q += 4 + if status >= 0xC0 { 1 } else { 0 } + packed;
        }
        if q > buf.len() {
            return None;
        }
    }
}

fn run<R: io::BufRead, W: io::Write>(fmt: c_int, opt: &LzbOptions, input: &mut R, output: &mut W) -> error::Result<()> {
    // The batch decoder wants the whole stream, the reference reads only as much of its BufRead as the stream needs.
    // 1) Decode what the reader's buffer shows (`fill_buf` does not consume).  For in-memory readers (`&[u8]`, `Cursor`)
    //    that is everything: the decode is final and exactly `consumed` bytes are taken, trailing bytes stay unread like
    //    in the reference.
    // 2) A streaming reader whose buffer ends inside the stream: LZMA2 is read up to the end its chunk headers give
    //    (exact for well-formed framing); .xz takes the rest of the reader (the reference rejects trailing data anyway,
    //    xz.rs:88-92); .lzma takes the rest of the reader too -- the one documented deviation: its compressed length
    //    is only known by decoding it.
    let first: Vec<u8> = input.fill_buf()?.to_vec(); // a copy: the borrow of `input` must end before consume()
    let (st, out, consumed) = decode_slice(fmt, opt, &first)?;
    let (st, out) = if !ran_out_of_input(&st) {
        // success, or an error the reference would raise on the same bytes: partial output still reaches the sink
        input.consume(consumed.min(first.len()));
        (st, out)
    } else {
        let mut buf = first;
        let n = buf.len();
        input.consume(n);
        if fmt == FMT_LZMA2 {
            loop {
                if let Some(end) = lzma2_extent(&buf) {
                    debug_assert!(end <= buf.len());
                    break;
                }
                let more: Vec<u8> = input.fill_buf()?.to_vec();
                if more.is_empty() {
                    break; // truncated stream: the decode below reports the reference's error
                }
                // take only up to the end of the stream if it lies inside this refill
                let have = buf.len();
                buf.extend_from_slice(&more);
                let take = match lzma2_extent(&buf) {
                    Some(end) => end - have,
                    None => more.len(),
                };
                buf.truncate(have + take);
                input.consume(take);
            }
        } else {
            input.read_to_end(&mut buf)?;
        }
        let (st, out, _consumed) = decode_slice(fmt, opt, &buf)?;
        (st, out)
    };
    // like the reference, partial output reaches the sink even when the stream then fails
    if !out.is_empty() {
        output.write_all(&out)?;
    }
    if st.code != 0 {
        return Err(to_error(&st));
    }
    output.flush()?;
    Ok(())
}

/// Batch form of an entry point: n independent streams in one launch (what `bench.py` drives through the same C call).
/// Element i is what `f(&mut inputs[i], &mut Vec::new())` of the one-shot function returns, with the decoded bytes.
fn decode_many(fmt: c_int, opt: &LzbOptions, inputs: &[&[u8]]) -> Vec<error::Result<Vec<u8>>> {
    let n = inputs.len();
    if n == 0 {
        return Vec::new();
    }
    let fail = |msg: String| -> Vec<error::Result<Vec<u8>>> {
        (0..n).map(|_| Err(error::Error::IoError(io::Error::new(io::ErrorKind::Other, msg.clone())))).collect()
    };
    let mut in_off = Vec::with_capacity(n + 1);
    let mut blob = Vec::with_capacity(inputs.iter().map(|s| s.len()).sum::<usize>() + 16);
    in_off.push(0u64);
    for s in inputs {
        blob.extend_from_slice(s);
        in_off.push(blob.len() as u64);
    }
    blob.extend_from_slice(&[0u8; 16]);
    let mut cap = vec![0u64; n];
    let rc = match with_ctx(|c| unsafe { lzb_scan(c, fmt, opt, blob.as_ptr(), in_off.as_ptr(), n as u32, cap.as_mut_ptr()) }) {
        Ok(rc) => rc,
        Err(e) => return fail(e.to_string()),
    };
    if rc != 0 {
        return fail(format!("lzb_scan failed: {}", rc));
    }
    let mut pending: Vec<usize> = (0..n).collect();
    let mut results: Vec<Option<error::Result<Vec<u8>>>> = (0..n).map(|_| None).collect();
    // streams whose size the headers do not give (end-marker .lzma) may report LZB_E_CAPACITY: rerun those, larger
    while !pending.is_empty() {
        let m = pending.len();
        let mut sub_in = Vec::with_capacity(m + 1);
        let mut sub_blob = Vec::new();
        let whole = m == n;
        if !whole {
            sub_in.push(0u64);
            for &i in &pending {
                sub_blob.extend_from_slice(inputs[i]);
                sub_in.push(sub_blob.len() as u64);
            }
            sub_blob.extend_from_slice(&[0u8; 16]);
        }
        let (src, src_off): (&[u8], &[u64]) = if whole { (&blob, &in_off) } else { (&sub_blob, &sub_in) };
        let mut out_off = vec![0u64; m + 1];
        for (k, &i) in pending.iter().enumerate() {
            out_off[k + 1] = out_off[k] + ((cap[i] + 15) & !15);
        }
        let mut out = vec![0u8; out_off[m] as usize + 16];
        let (mut out_len, mut consumed, mut st) = (vec![0u64; m], vec![0u64; m], vec![LzbStatus::default(); m]);
        let rc = with_engine(|e| unsafe {
            if e.multi.is_null() {
                lzb_decode_batch(e.one, fmt, opt, src.as_ptr(), src_off.as_ptr(), m as u32, out.as_mut_ptr(), out_off.as_ptr(),
                                 out_len.as_mut_ptr(), consumed.as_mut_ptr(), st.as_mut_ptr())
            } else {
                lzb_decode_batch_multi(e.multi, fmt, opt, src.as_ptr(), src_off.as_ptr(), m as u32, out.as_mut_ptr(),
                                       out_off.as_ptr(), out_len.as_mut_ptr(), consumed.as_mut_ptr(), st.as_mut_ptr(),
                                       std::ptr::null_mut())
            }
        });
        match rc {
            Ok(0) => {}
            Ok(rc) => return fail(format!("lzb_decode_batch failed: {}", rc)),
            Err(e) => return fail(e.to_string()),
        }
        let mut again = Vec::new();
        for (k, &i) in pending.iter().enumerate() {
            if st[k].code == -1 && cap[i] < 0xFFFF_F000 {
                cap[i] = (cap[i] * 2).max(st[k].a0 + 65536).min(0xFFFF_F000);
                again.push(i);
                continue;
            }
            results[i] = Some(if st[k].code == 0 {
                Ok(out[out_off[k] as usize..(out_off[k] + out_len[k]) as usize].to_vec())
            } else {
                Err(to_error(&st[k]))
            });
        }
        pending = again;
    }
    results.into_iter().map(|r| r.unwrap()).collect()
}
/// `lzma_decompress` over a batch of independent `.lzma` streams.
pub fn lzma_decompress_batch(inputs: &[&[u8]]) -> Vec<error::Result<Vec<u8>>> {
    decode_many(FMT_LZMA, &LzbOptions::default(), inputs)
}
/// `lzma_decompress_with_options` over a batch.
pub fn lzma_decompress_batch_with_options(inputs: &[&[u8]], opts: &decompress::Options) -> Vec<error::Result<Vec<u8>>> {
    decode_many(FMT_LZMA, &options(opts), inputs)
}
/// `lzma2_decompress` over a batch of independent raw LZMA2 streams (the benchmarked call).
pub fn lzma2_decompress_batch(inputs: &[&[u8]]) -> Vec<error::Result<Vec<u8>>> {
    decode_many(FMT_LZMA2, &LzbOptions::default(), inputs)
}
/// `xz_decompress` over a batch of `.xz` files.
pub fn xz_decompress_batch(inputs: &[&[u8]]) -> Vec<error::Result<Vec<u8>>> {
    decode_many(FMT_XZ, &LzbOptions::default(), inputs)
}

fn options(o: &decompress::Options) -> LzbOptions {
    let mut r = LzbOptions::default();
    match o.unpacked_size {
        decompress::UnpackedSize::ReadFromHeader => r.unpacked_mode = 0,
        decompress::UnpackedSize::ReadHeaderButUseProvided(x) => {
            r.unpacked_mode = 1;
            r.has_provided = x.is_some() as u8;
            r.provided = x.unwrap_or(0);
        }
        decompress::UnpackedSize::UseProvided(x) => {
            r.unpacked_mode = 2;
            r.has_provided = x.is_some() as u8;
            r.provided = x.unwrap_or(0);
        }
    }
    if let Some(m) = o.memlimit {
        r.has_memlimit = 1;
        r.memlimit = m as u64;
    }
    r
}

/// `lzma_rs::decompress::Stream` (feature `stream`, src/decode/stream.rs:66-346) as a façade over the batch path:
/// `write` buffers (and still rejects an invalid properties byte at once, stream.rs:157-190), `finish` decodes the
/// whole stream on the GPU.  Same results as the reference; data errors surface in `finish` instead of `write`.
/// With `allow_incomplete` on an unknown-size stream this shim returns every byte of every complete symbol, which can be
/// a few bytes MORE than the reference (it stops as soon as its input is exhausted); the Python façade
/// (`lzma_rs_b200.Stream._incomplete_output`) shows the two extra decodes that trim the difference.
pub struct Stream<W: io::Write> {
    output: Option<W>,
    buf: Vec<u8>,
    options: decompress::Options,
}
impl<W: io::Write> Stream<W> {
    pub fn new(output: W) -> Self { Self::new_with_options(&decompress::Options::default(), output) }
    pub fn new_with_options(options: &decompress::Options, output: W) -> Self {
        Stream { output: Some(output), buf: Vec::new(), options: *options }
    }
    pub fn get_output(&self) -> Option<&W> { self.output.as_ref() }
    pub fn get_output_mut(&mut self) -> Option<&mut W> { self.output.as_mut() }
    pub fn finish(mut self) -> error::Result<W> {
        let mut out = self.output.take().ok_or_else(|| {
            error::Error::LzmaError("can't finish stream because of previous write error".to_string())
        })?;
        if self.buf.is_empty() { return Ok(out); }
        let hdr = if let decompress::UnpackedSize::UseProvided(_) = self.options.unpacked_size { 5 } else { 13 };
        if self.buf.len() < hdr + 5 { return Err(error::Error::LzmaError("failed to read header".to_string())); }
        let mut o = options(&self.options);
        o.allow_incomplete = self.options.allow_incomplete as u8; // a stream-API option: the one-shot functions ignore it
        run(FMT_LZMA, &o, &mut &self.buf[..], &mut out)?;
        Ok(out)
    }
}
impl<W: io::Write> io::Write for Stream<W> {
    fn write(&mut self, data: &[u8]) -> io::Result<usize> {
        if self.output.is_none() { return Ok(0); }
        let first = self.buf.is_empty();
        self.buf.extend_from_slice(data);
        if first && !self.buf.is_empty() && self.buf[0] >= 225 {
            self.output = None;
            return Err(io::Error::new(io::ErrorKind::Other,
                format!("LZMA header invalid properties: {} must be < 225", self.buf[0])));
        }
        Ok(data.len())
    }
    fn flush(&mut self) -> io::Result<()> { match self.output.as_mut() { Some(o) => o.flush(), None => Ok(()) } }
}

/// `lzma_rs::lzma_decompress` (src/lib.rs:44-49)
pub fn lzma_decompress<R: io::BufRead, W: io::Write>(input: &mut R, output: &mut W) -> error::Result<()> {
    lzma_decompress_with_options(input, output, &decompress::Options::default())
}
/// `lzma_rs::lzma_decompress_with_options` (src/lib.rs:52-60)
pub fn lzma_decompress_with_options<R: io::BufRead, W: io::Write>(
    input: &mut R, output: &mut W, opts: &decompress::Options,
) -> error::Result<()> {
    run(FMT_LZMA, &options(opts), input, output)
}
/// `lzma_rs::compress` (src/encode/options.rs:1-30)
pub mod compress {
    #[derive(Clone, Copy, Debug)]
    pub enum UnpackedSize {
        WriteToHeader(Option<u64>),
        SkipWritingToHeader,
    }
    impl Default for UnpackedSize {
        fn default() -> UnpackedSize { UnpackedSize::WriteToHeader(None) }
    }
    #[derive(Clone, Copy, Debug, Default)]
    pub struct Options {
        pub unpacked_size: UnpackedSize,
    }
}

// The reference's encoders are format writers (literals only / stored chunks only); the GPU writes the same bytes.
fn encode<R: io::BufRead, W: io::Write>(fmt: c_int, opt: &LzbCompressOptions, input: &mut R, output: &mut W) -> io::Result<()> {
    let mut buf = Vec::new();
    input.read_to_end(&mut buf)?;
    let mut cap = unsafe { lzb_encode_bound(fmt, opt, buf.len() as u64) };
    loop {
        let mut out = vec![0u8; cap as usize + 16];
        let (in_off, out_off) = ([0u64, buf.len() as u64], [0u64, cap]);
        let (mut out_len, mut st) = (0u64, LzbStatus::default());
        let rc = with_ctx(|c| unsafe {
            lzb_encode_batch(c, fmt, opt, buf.as_ptr(), in_off.as_ptr(), 1, out.as_mut_ptr(), out_off.as_ptr(), &mut out_len, &mut st)
        })?;
        if rc != 0 {
            return Err(io::Error::new(io::ErrorKind::Other, format!("lzma_b200 call failed: {}", rc)));
        }
        if st.code == -1 { // LZB_E_CAPACITY: a0 = bytes needed (adversarial input for the literal coder)
            cap = st.a0;
            continue;
        }
        return output.write_all(&out[..out_len as usize]);
    }
}
/// `lzma_rs::lzma_compress` (src/lib.rs:63-69)
pub fn lzma_compress<R: io::BufRead, W: io::Write>(input: &mut R, output: &mut W) -> io::Result<()> {
    lzma_compress_with_options(input, output, &compress::Options::default())
}
/// `lzma_rs::lzma_compress_with_options` (src/lib.rs:72-80)
pub fn lzma_compress_with_options<R: io::BufRead, W: io::Write>(
    input: &mut R, output: &mut W, options: &compress::Options,
) -> io::Result<()> {
    let mut o = LzbCompressOptions::default();
    match options.unpacked_size {
        compress::UnpackedSize::SkipWritingToHeader => o.skip_size_field = 1,
        compress::UnpackedSize::WriteToHeader(Some(x)) => { o.has_value = 1; o.value = x; }
        compress::UnpackedSize::WriteToHeader(None) => {}
    }
    encode(FMT_LZMA, &o, input, output)
}
/// `lzma_rs::lzma2_compress` (src/lib.rs:91-97)
pub fn lzma2_compress<R: io::BufRead, W: io::Write>(input: &mut R, output: &mut W) -> io::Result<()> {
    encode(FMT_LZMA2, &LzbCompressOptions::default(), input, output)
}
/// `lzma_rs::xz_compress` (src/lib.rs:108-110)
pub fn xz_compress<R: io::BufRead, W: io::Write>(input: &mut R, output: &mut W) -> io::Result<()> {
    encode(FMT_XZ, &LzbCompressOptions::default(), input, output)
}

/// `lzma_rs::decompress::raw` (feature `raw_decoder`, src/lib.rs:29-35) over `lzb_raw_*`: decoder objects whose
/// DecoderState -- probabilities, state, rep distances, and for LZMA2 the properties of the last props reset -- lives on
/// the device and survives from one `decompress` to the next until `reset`, like the reference's
/// (src/decode/lzma.rs:597-648, src/decode/lzma2.rs:11-82); every call starts an empty output window.
pub mod raw {
    use super::*;
    /// src/decode/lzma.rs:41-66
    #[derive(Debug, Copy, Clone)]
    pub struct LzmaProperties { pub lc: u32, pub lp: u32, pub pb: u32 }
    /// src/decode/lzma.rs:68-93
    #[derive(Debug, Copy, Clone)]
    pub struct LzmaParams { properties: LzmaProperties, dict_size: u32, unpacked_size: Option<u64> }
    impl LzmaParams {
        pub fn new(properties: LzmaProperties, dict_size: u32, unpacked_size: Option<u64>) -> LzmaParams {
            LzmaParams { properties, dict_size, unpacked_size }
        }
    }
    #[repr(C)]
    pub(crate) struct LzbRaw { _private: [u8; 0] }
    extern "C" {
        fn lzb_raw_create(ctx: *mut LzbCtx, fmt: c_int, lc: u32, lp: u32, pb: u32, dict_size: u32, raw: *mut *mut LzbRaw) -> c_int;
        fn lzb_raw_reset(raw: *mut LzbRaw) -> c_int;
        fn lzb_raw_decompress(
            raw: *mut LzbRaw, opt: *const LzbOptions, input: *const u8, in_len: usize, out: *mut *mut u8,
            out_len: *mut usize, consumed: *mut usize, st: *mut LzbStatus,
        ) -> c_int;
        fn lzb_raw_destroy(raw: *mut LzbRaw);
    }
    struct Handle(*mut LzbRaw);
    unsafe impl Send for Handle {}
    impl Drop for Handle {
        fn drop(&mut self) { unsafe { lzb_raw_destroy(self.0) } }
    }
    fn create(fmt: c_int, lc: u32, lp: u32, pb: u32, dict_size: u32) -> error::Result<Handle> {
        let mut h: *mut LzbRaw = std::ptr::null_mut();
        let rc = with_ctx(|c| unsafe { lzb_raw_create(c, fmt, lc, lp, pb, dict_size, &mut h) })?;
        if rc != 0 {
            return Err(error::Error::IoError(io::Error::new(io::ErrorKind::Other, format!("lzb_raw_create failed: {}", rc))));
        }
        Ok(Handle(h))
    }
    /// One decode through the decoder object.  The raw decoders take a reader like the one-shot functions; in-memory
    /// readers keep their trailing bytes unread (exactly `consumed` bytes are taken).
    fn run_raw<R: io::BufRead, W: io::Write>(h: &Handle, opt: &LzbOptions, input: &mut R, output: &mut W) -> error::Result<()> {
        let first: Vec<u8> = input.fill_buf()?.to_vec();
        let n = first.len();
        let mut buf = first;
        let mut whole = false;
        loop {
            let mut out: *mut u8 = std::ptr::null_mut();
            let (mut out_len, mut consumed) = (0usize, 0usize);
            let mut st = LzbStatus::default();
            let rc = unsafe { lzb_raw_decompress(h.0, opt, buf.as_ptr(), buf.len(), &mut out, &mut out_len, &mut consumed, &mut st) };
            if rc != 0 {
                return Err(error::Error::IoError(io::Error::new(io::ErrorKind::Other, format!("lzma_b200 call failed: {}", rc))));
            }
            let data = if out_len > 0 { unsafe { std::slice::from_raw_parts(out, out_len) }.to_vec() } else { Vec::new() };
            unsafe { lzb_free(out as *mut c_void) };
            if ran_out_of_input(&st) && !whole {
                // a streaming reader whose buffer ended inside the stream: a failed call commits nothing to the decoder's
                // state, so decode again with the whole reader (documented: INTEGRATION.md section 3)
                input.consume(n);
                input.read_to_end(&mut buf)?;
                whole = true;
                continue;
            }
            if !whole {
                input.consume(consumed.min(n));
            }
            if !data.is_empty() {
                output.write_all(&data)?;
            }
            if st.code != 0 {
                return Err(to_error(&st));
            }
            output.flush()?;
            return Ok(());
        }
    }
    /// src/decode/lzma.rs:597-648
    pub struct LzmaDecoder { params: LzmaParams, memlimit: Option<usize>, handle: Handle }
    impl LzmaDecoder {
        pub fn new(params: LzmaParams, memlimit: Option<usize>) -> error::Result<LzmaDecoder> {
            let p = params.properties;
            assert!(p.lc <= 8 && p.lp <= 4 && p.pb <= 4);
            let handle = create(FMT_LZMA, p.lc, p.lp, p.pb, params.dict_size)?;
            Ok(LzmaDecoder { params, memlimit, handle })
        }
        /// src/decode/lzma.rs:620-627
        pub fn reset(&mut self, unpacked_size: Option<Option<u64>>) {
            if let Some(u) = unpacked_size { self.params.unpacked_size = u; }
            unsafe { lzb_raw_reset(self.handle.0) };
        }
        pub fn decompress<W: io::Write, R: io::BufRead>(&mut self, input: &mut R, output: &mut W) -> error::Result<()> {
            let opts = decompress::Options {
                unpacked_size: decompress::UnpackedSize::UseProvided(self.params.unpacked_size),
                memlimit: self.memlimit, allow_incomplete: false,
            };
            run_raw(&self.handle, &options(&opts), input, output)
        }
    }
    /// src/decode/lzma2.rs:11-82
    pub struct Lzma2Decoder { handle: Handle }
    impl Lzma2Decoder {
        pub fn new() -> Lzma2Decoder {
            Lzma2Decoder { handle: create(FMT_LZMA2, 0, 0, 0, 0).expect("lzb_raw_create") }
        }
        pub fn reset(&mut self) { unsafe { lzb_raw_reset(self.handle.0) }; }
        pub fn decompress<W: io::Write, R: io::BufRead>(&mut self, input: &mut R, output: &mut W) -> error::Result<()> {
            run_raw(&self.handle, &LzbOptions::default(), input, output)
        }
    }
    impl Default for Lzma2Decoder {
        fn default() -> Self { Self::new() }
    }
}

/// `lzma_rs::lzma2_decompress` (src/lib.rs:83-88)
pub fn lzma2_decompress<R: io::BufRead, W: io::Write>(input: &mut R, output: &mut W) -> error::Result<()> {
    run(FMT_LZMA2, &LzbOptions::default(), input, output)
}
/// `lzma_rs::xz_decompress` (src/lib.rs:100-105)
pub fn xz_decompress<R: io::BufRead, W: io::Write>(input: &mut R, output: &mut W) -> error::Result<()> {
    run(FMT_XZ, &LzbOptions::default(), input, output)
}
