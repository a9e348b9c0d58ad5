/*
 * lzma_b200.h -- C ABI of the B200-native many-stream LZMA / LZMA2 / XZ decoder.
 *
 * This is the drop-in boundary for the decode path of gendx/lzma-rs (reference @ 1f14478).
 * The reference has NO FFI/plugin interface of its own: its boundary is three generic Rust
 * free functions over io::BufRead / io::Write,
 *
 *     lzma_rs::lzma_decompress(input, output)                       src/lib.rs:44-49
 *     lzma_rs::lzma_decompress_with_options(input, output, options) src/lib.rs:52-60
 *     lzma_rs::lzma2_decompress(input, output)                      src/lib.rs:83-88
 *     lzma_rs::xz_decompress(input, output)                         src/lib.rs:100-105
 *
 * so every entry point below is what a Rust shim for those functions binds (see
 * INTEGRATION.md and rust/lzma_b200/src/lib.rs): plain pointers and sizes, no C++ or torch
 * types, never throws, never aborts, no callbacks.  Exported by
 * lzma_rs_b200/liblzma_b200.so (sm_100a only; there is no CPU fallback -- every stream is
 * decoded by the CUDA kernels, and every call fails with LZB_RC_NO_DEVICE without a GPU).
 */
#ifndef LZMA_B200_H
#define LZMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZB_ABI_VERSION 2

/* ---- call-level return codes (infrastructure; distinct from per-stream decode status) ---- */
enum {
    LZB_RC_OK = 0,
    LZB_RC_BAD_ARG = -1,
    LZB_RC_NO_DEVICE = -2, /* no CUDA device / driver: the product has no CPU path */
    LZB_RC_CUDA = -3,      /* a CUDA call failed; see lzb_last_error() */
    LZB_RC_OOM = -4
};

/* ---- stream formats = the three reference entry points ---- */
typedef enum lzb_format {
    LZB_FMT_LZMA = 0,  /* .lzma "alone": lzma_decompress[_with_options], lib.rs:44-60 */
    LZB_FMT_LZMA2 = 1, /* raw LZMA2:     lzma2_decompress,               lib.rs:83-88 */
    LZB_FMT_XZ = 2     /* .xz:           xz_decompress,                   lib.rs:100-105 */
} lzb_format;

/* ---- decompress::Options (src/decode/options.rs:3-43), passed by pointer; NULL = default ---- */
enum { LZB_UNPACKED_READ_FROM_HEADER = 0, LZB_UNPACKED_READ_HEADER_BUT_USE_PROVIDED = 1, LZB_UNPACKED_USE_PROVIDED = 2 };
typedef struct lzb_options {
    uint8_t unpacked_mode; /* UnpackedSize variant */
    uint8_t has_provided;  /* the variant's Option<u64> is Some */
    uint8_t has_memlimit;  /* memlimit: Option<usize> is Some */
    uint8_t allow_incomplete; /* Options::allow_incomplete (the stream API's flag, stream.rs:136-147): a .lzma stream whose
                               * input ends in the middle of a symbol, or whose decoded length disagrees with its declared
                               * size, is not an error; every byte decoded from complete symbols is returned (LZB_OK).
                               * For unknown-size streams that is a few symbols MORE than the reference's incremental
                               * decoder returns (it stops as soon as its input is exhausted, lzma.rs:450-452; the batch
                               * decoder also takes the symbols behind that point that need no further byte); the
                               * Stream facades of the host layers trim the difference. */
    uint8_t reserved[4];
    uint64_t provided;
    uint64_t memlimit;
} lzb_options;

/* ---- per-stream status: one code per error site of the reference.
 * `kind` is the error::Error variant (src/error.rs:7-17); a0..a2 are the values the reference
 * formats into its message.  lzb_format_error() renders the reference's exact Display string. */
enum { LZB_KIND_OK = 0, LZB_KIND_IO = 1, LZB_KIND_HEADER_TOO_SHORT = 2, LZB_KIND_LZMA = 3, LZB_KIND_XZ = 4,
       LZB_KIND_INTERNAL = 5 /* not a reference error: capacity / unsupported on the GPU path */ };

enum {
    LZB_OK = 0,
    LZB_E_IO_EOF = 1,               /* IoError(UnexpectedEof): rangecoder.rs:64, xz.rs read_u8/read_u32 `?` sites */
    LZB_E_HEADER_TOO_SHORT = 2,     /* lzma.rs:101,119-121,133-135,144-146 */
    LZB_E_LZMA_PROPS = 3,           /* lzma.rs:104-109   a0=props */
    LZB_E_LZMA_STREAM_TOO_SHORT = 4,/* lzma.rs:643-644 */
    LZB_E_EOS_MORE_BYTES = 5,       /* lzma.rs:378-380 */
    LZB_E_UNPACKED_MISMATCH = 6,    /* lzma.rs:513-521   a0=expected a1=got */
    LZB_E_MATCH_DIST_DICT = 7,      /* lzbuffer.rs:241-246 a0=dist a1=dict_size */
    LZB_E_MATCH_DIST_OUT = 8,       /* lzbuffer.rs:100-105,247-252 a0=dist a1=len */
    LZB_E_LZ_DIST_DICT = 9,         /* lzbuffer.rs:274-279 a0=dist a1=dict_size */
    LZB_E_LZ_DIST_OUT = 10,         /* lzbuffer.rs:128-133,280-285 a0=dist a1=len */
    LZB_E_MEMLIMIT = 11,            /* lzbuffer.rs:213-217 a0=memlimit */
    LZB_E_L2_STATUS_EOF = 12,       /* lzma2.rs:60-62 */
    LZB_E_L2_INVALID_STATUS = 13,   /* lzma2.rs:94-99    a0=status */
    LZB_E_L2_UNPACKED_EOF = 14,     /* lzma2.rs:128-130, 204-206 */
    LZB_E_L2_PACKED_EOF = 15,       /* lzma2.rs:133-135 */
    LZB_E_L2_PROPS_EOF = 16,        /* lzma2.rs:153-155 */
    LZB_E_L2_PROPS_RANGE = 17,      /* lzma2.rs:158-163  a0=props */
    LZB_E_L2_PROPS_LCLP = 18,       /* lzma2.rs:170-175  a0=lc a1=lp */
    LZB_E_L2_STORED_EOF = 19,       /* lzma2.rs:220-225  a0=size */
    LZB_E_L2_INPUT_TOO_SHORT = 20,  /* lzma2.rs:190-191 */
    /* XZ container (host-side walk of xz.rs / xz/header.rs / xz/mod.rs) */
    LZB_E_XZ_MAGIC = 30,            /* header.rs:24-29 */
    LZB_E_XZ_HEADER_CRC = 31,       /* header.rs:38-44, xz.rs:217-224  a0=read a1=computed */
    LZB_E_XZ_FLAGS_NULL = 32,       /* xz/mod.rs:27-32   a0=byte */
    LZB_E_XZ_CHECK_METHOD = 33,     /* xz/mod.rs:65-75   a0=id */
    LZB_E_XZ_BLOCK_FLAGS = 34,      /* xz.rs:375-380     a0=flags */
    LZB_E_XZ_FILTER_ID = 35,        /* xz.rs:178-183     a0=id */
    LZB_E_XZ_PROPS_SIZE = 36,       /* xz.rs:412-417     a0=size a1=header_size */
    LZB_E_XZ_PROPS_READ = 37,       /* xz.rs:421-426     a0=size */
    LZB_E_XZ_HEADER_PADDING = 38,   /* xz.rs:435-439 */
    LZB_E_XZ_FILTER_PROPS = 39,     /* xz.rs:343-348 */
    LZB_E_XZ_PACKED_SIZE = 40,      /* xz.rs:232-238     a0=expected a1=got */
    LZB_E_XZ_UNPACKED_SIZE = 41,    /* xz.rs:255-261     a0=expected a1=got */
    LZB_E_XZ_BLOCK_PADDING = 42,    /* xz.rs:272-278 */
    LZB_E_XZ_BLOCK_CRC32 = 43,      /* xz.rs:305-312     a0=read a1=computed */
    LZB_E_XZ_BLOCK_CRC64 = 44,      /* xz.rs:316-323     a0=read a1=computed */
    LZB_E_XZ_SHA256 = 45,           /* xz.rs:326-330 */
    LZB_E_XZ_INDEX_COUNT = 46,      /* xz.rs:109-116     a0=index a1=records */
    LZB_E_XZ_INDEX_UNPADDED = 47,   /* xz.rs:121-127     a0=record a1=actual a2=index */
    LZB_E_XZ_INDEX_UNPACKED = 48,   /* xz.rs:129-135     a0=record a1=actual a2=index */
    LZB_E_XZ_INDEX_PADDING = 49,    /* xz.rs:150-155 */
    LZB_E_XZ_INDEX_CRC = 50,        /* xz.rs:162-168     a0=read a1=computed */
    LZB_E_XZ_MULTIBYTE = 51,        /* xz.rs:461-463 */
    LZB_E_XZ_INDEX_SIZE = 52,       /* xz.rs:52-58       a0=expected a1=got */
    LZB_E_XZ_FLAGS_MISMATCH = 53,   /* xz.rs:65-70       a0=header check a1=footer check */
    LZB_E_XZ_FOOTER_CRC = 54,       /* xz.rs:73-79       a0=read a1=computed */
    LZB_E_XZ_FOOTER_MAGIC = 55,     /* xz.rs:81-86 */
    LZB_E_XZ_TRAILING_DATA = 56,    /* xz.rs:88-92 */
    /* not reference errors (kind = LZB_KIND_INTERNAL) */
    LZB_E_CAPACITY = -1,    /* caller's output capacity too small; a0 = bytes needed (lower bound if the size is unknown) */
    LZB_E_UNSUPPORTED = -2, /* outside the GPU path's limits (stream or output >= 4 GiB - 4 KiB) */
    LZB_E_INPUT_TIMEOUT = -3 /* host API: the stream's input never reached the device (lost upload); nothing decoded */
};

typedef struct lzb_status {
    int32_t code; /* LZB_OK / LZB_E_* */
    int32_t kind; /* LZB_KIND_* */
    uint64_t a0, a1, a2;
} lzb_status;

typedef struct lzb_ctx lzb_ctx;

/* One context per CUDA device (one process per GPU; multi-GPU = one ctx per rank).
 * device < 0 -> the current device. */
int lzb_create(lzb_ctx **ctx, int device);
void lzb_destroy(lzb_ctx *ctx);
/* Text of the last infrastructure failure on this ctx (CUDA error string); "" if none. */
const char *lzb_last_error(const lzb_ctx *ctx);
int lzb_abi_version(void);

/* Host-side size scan (no GPU work): walks .lzma headers / LZMA2 chunk headers / the XZ container
 * (lzma.rs:96-161, lzma2.rs:128-136,204-207, xz.rs:356-446) of stream i = in[in_off[i], in_off[i+1])
 * and writes the output capacity lzb_decode_batch needs for it: exact size for well-formed LZMA2 /
 * XZ / known-size .lzma, a heuristic bound for end-marker .lzma (decode reports LZB_E_CAPACITY with
 * the bytes needed if it was too small).  st may be NULL. */
int lzb_scan(lzb_ctx *ctx, int fmt, const lzb_options *opt, const uint8_t *in, const uint64_t *in_off,
             uint32_t n, uint64_t *capacity);

/* Decode n independent streams from HOST memory into HOST memory (the reference-facing path):
 *   stream i input  = in[in_off[i], in_off[i+1])
 *   stream i output -> out[out_off[i], ...), capacity out_off[i+1]-out_off[i]
 *   out_len[i]  = bytes the reference would have written to its io::Write sink (also on error:
 *                 the reference's partial output, e.g. whole validated XZ blocks)
 *   consumed[i] = bytes the reference would have consumed from its io::BufRead (on success)
 *   st[i]       = per-stream status
 * H2D copy, kernels and D2H copy all happen inside the call (pinned buffers copy at full PCIe rate).
 * Returns LZB_RC_*; per-stream failures do not fail the call. */
int lzb_decode_batch(lzb_ctx *ctx, int fmt, const lzb_options *opt, const uint8_t *in, const uint64_t *in_off,
                     uint32_t n, uint8_t *out, const uint64_t *out_off, uint64_t *out_len, uint64_t *consumed,
                     lzb_status *st);

/* Same for DEVICE-resident raw streams (fmt = LZB_FMT_LZMA2, or LZB_FMT_LZMA): d_in / d_out are device
 * pointers (d_in readable up to the next multiple of 4 bytes); offsets and result arrays are host
 * arrays.  Runs on `cuda_stream` (a cudaStream_t passed as void*, NULL = the ctx's own stream) and
 * synchronises it before returning the per-stream results.  This is what bench.py times as the
 * kernel-only `value`. */
int lzb_decode_batch_device(lzb_ctx *ctx, int fmt, const lzb_options *opt, const uint8_t *d_in,
                            const uint64_t *in_off, uint32_t n, uint8_t *d_out, const uint64_t *out_off,
                            uint64_t *out_len, uint64_t *consumed, lzb_status *st, void *cuda_stream);

/* Two-phase variant of the device path for timing and overlap: prepare uploads the offsets and runs the
 * per-stream scan kernel (header parse / LZMA2 framing walk) once; launch enqueues ONLY the decode
 * kernel on `cuda_stream` (no sync); collect synchronises and fetches results.  A prepared batch can
 * be launched repeatedly.  d_in and d_out must be 16-byte aligned.  Every batch owns its device workspaces: different
 * batches of one context may be in flight on different streams at the same time; launches of the SAME batch must not
 * overlap (they share its result and workspace buffers).  The host entry points (lzb_decode_batch, lzb_encode_batch,
 * lzb_decompress_alloc) serialise on the context. */
typedef struct lzb_batch lzb_batch;
int lzb_batch_prepare(lzb_ctx *ctx, int fmt, const lzb_options *opt, const uint8_t *d_in, const uint64_t *in_off,
                      uint32_t n, uint8_t *d_out, const uint64_t *out_off, lzb_batch **batch);
int lzb_batch_launch(lzb_batch *batch, void *cuda_stream);
int lzb_batch_collect(lzb_batch *batch, void *cuda_stream, uint64_t *out_len, uint64_t *consumed, lzb_status *st);
void lzb_batch_destroy(lzb_batch *batch);
/* Number of kernels one lzb_batch_launch enqueues (for bench.py's gpu_launches claim). */
int lzb_batch_kernels_per_launch(const lzb_batch *batch);
/* Name of the decode kernel lzb_batch_launch runs first for this batch (the variant the planner picked: throughput /
 * placement-plan / fill / copy / latency form, or the stored-chunk copy kernel); static string.  Measurement aid. */
const char *lzb_batch_kernel_name(const lzb_batch *batch);

/* ---- device-side sizing and layout (replaces the reference's incremental growth of its output
 * Vec: the sizes the reference learns chunk by chunk in lzma2.rs:128-136, 204-207 / from the .lzma header in
 * lzma.rs:128-148 are computed for the whole batch on the device, and so is the output layout).  Everything is a
 * DEVICE pointer except `total`: a batch that was produced on the GPU (or arrived over NVLink) is sized, laid out and
 * decoded without its offsets or sizes ever visiting the host.
 *   d_in_off   n+1 offsets of the streams in d_in
 *   d_capacity n   output bytes stream i needs (same rule as lzb_scan for raw formats), may be NULL
 *   d_out_off  n+1 exclusive prefix sum of the capacities rounded up to 16 bytes, may be NULL
 *   total      host, *total = d_out_off[n] (the output blob to allocate), may be NULL (then no synchronisation happens)
 * fmt = LZB_FMT_LZMA or LZB_FMT_LZMA2.  Runs on `cuda_stream` (NULL = the ctx's stream). */
int lzb_scan_device(lzb_ctx *ctx, int fmt, const lzb_options *opt, const uint8_t *d_in, const uint64_t *d_in_off,
                    uint32_t n, uint64_t *d_capacity, uint64_t *d_out_off, uint64_t *total, void *cuda_stream);
/* lzb_batch_prepare with DEVICE offset arrays (e.g. the d_out_off lzb_scan_device produced). */
int lzb_batch_prepare_device(lzb_ctx *ctx, int fmt, const lzb_options *opt, const uint8_t *d_in, const uint64_t *d_in_off,
                             uint32_t n, uint8_t *d_out, const uint64_t *d_out_off, lzb_batch **batch);

/* ---- multi-GPU (lzb_create(ctx**, dev_ids, n_dev); 8(e): independent streams shard with no data-path
 * collective).  Two forms:
 *  (1) one process, several devices: lzb_multi owns one lzb_ctx per device and lzb_decode_batch_multi splits a HOST
 *      batch into contiguous stream ranges balanced by compressed bytes; every device uploads its own range over its
 *      own PCIe link, decodes, and streams its output pages back into the caller's buffer -- n_dev copies of
 *      lzb_decode_batch running concurrently.
 *  (2) one process per device (torchrun / MPI ranks): the batch lives in ONE rank's HBM; that rank exports its input
 *      and output blobs (lzb_ipc_export), the others open them (lzb_ipc_open: CUDA IPC, NVLink peer mapping) and call
 *      lzb_decode_batch_peer on their stream range: the compressed bytes are pulled over NVLink in chunks behind K1's
 *      input gate while the kernel already decodes, and K1 writes finished output pages straight into the owner's blob
 *      (peer stores), so scatter, decode and gather are one launch with no staging copy and no collective. */
typedef struct lzb_multi lzb_multi;
int lzb_create_multi(lzb_multi **m, const int *dev_ids, int n_dev); /* dev_ids NULL / n_dev <= 0: every visible device */
void lzb_destroy_multi(lzb_multi *m);
int lzb_multi_device_count(const lzb_multi *m);
lzb_ctx *lzb_multi_ctx(lzb_multi *m, int k); /* the k-th device's context (owned by m) */
const char *lzb_multi_last_error(const lzb_multi *m);
/* Same contract as lzb_decode_batch.  split[0..n_dev] (may be NULL) receives the stream ranges the devices got. */
int lzb_decode_batch_multi(lzb_multi *m, int fmt, const lzb_options *opt, const uint8_t *in, const uint64_t *in_off,
                           uint32_t n, uint8_t *out, const uint64_t *out_off, uint64_t *out_len, uint64_t *consumed,
                           lzb_status *st, uint32_t *split);

typedef struct lzb_ipc_handle {
    uint8_t handle[64]; /* cudaIpcMemHandle_t of the allocation that contains the pointer */
    uint64_t offset;    /* of the pointer inside that allocation */
    uint64_t bytes;     /* informational */
} lzb_ipc_handle;
int lzb_ipc_export(lzb_ctx *ctx, const void *d_ptr, uint64_t bytes, lzb_ipc_handle *h);
int lzb_ipc_open(lzb_ctx *ctx, const lzb_ipc_handle *h, void **d_ptr); /* maps the exporter's memory into this process */
int lzb_ipc_close(lzb_ctx *ctx, void *d_ptr);                          /* d_ptr as returned by lzb_ipc_open */
/* Decode streams whose bytes live in memory this device can reach but that is not its own HBM -- a peer GPU's blob
 * (lzb_ipc_open, or cudaDeviceEnablePeerAccess in one process) or pinned host memory -- into a blob of the same kind:
 *   stream i input = src_in[in_off[i], in_off[i+1]),  output -> dst_out[out_off[i], ...)   (out_off[i] 16-byte aligned)
 * fmt = LZB_FMT_LZMA or LZB_FMT_LZMA2; offsets / results are host arrays.  The framing scan (K2) reads src_in in
 * place; the input follows in chunks behind the gate; output pages are stored to dst_out by the decode kernel.
 * Synchronises before returning. */
int lzb_decode_batch_peer(lzb_ctx *ctx, int fmt, const lzb_options *opt, const uint8_t *src_in, const uint64_t *in_off,
                          uint32_t n, uint8_t *dst_out, const uint64_t *out_off, uint64_t *out_len, uint64_t *consumed,
                          lzb_status *st);

/* ---- decompress::raw::{LzmaDecoder, Lzma2Decoder} (feature raw_decoder; src/decode/lzma.rs:597-648,
 * src/decode/lzma2.rs:11-82): decoder OBJECTS whose DecoderState -- probabilities, state, rep[4], and for LZMA2 the
 * properties of the last props reset -- survives from one decompress() call to the next until reset() is called; the
 * output window does not (every call starts an empty one), exactly as in the reference.  The state lives in device
 * memory; a call runs on a copy and commits it when the decode is over (so a capacity retry starts from the same state).
 *   fmt = LZB_FMT_LZMA : lc/lp/pb/dict_size = LzmaParams (headerless payloads); the unpacked size of each call comes
 *                        from opt (has_provided / provided = Option<u64>, as LzmaParams::unpacked_size / reset(Some(..)))
 *   fmt = LZB_FMT_LZMA2: lc/lp/pb/dict_size ignored (DecoderState::new(0, 0, 0), lzma2.rs:24-34)
 * After a call that ended in a decode error the reference's state is whatever the failing symbol left behind (its tests
 * never reuse such a decoder); here a failed call commits nothing: the decoder is as it was before the call. */
typedef struct lzb_raw lzb_raw;
int lzb_raw_create(lzb_ctx *ctx, int fmt, uint32_t lc, uint32_t lp, uint32_t pb, uint32_t dict_size, lzb_raw **raw);
int lzb_raw_reset(lzb_raw *raw); /* DecoderState::reset_state (lzma.rs:216-249) */
int lzb_raw_decompress(lzb_raw *raw, const lzb_options *opt, const uint8_t *in, size_t in_len, uint8_t **out,
                       size_t *out_len, size_t *consumed, lzb_status *st);
void lzb_raw_destroy(lzb_raw *raw);

/* Single-stream convenience for the Rust/C++ shim: scan + decode + (for end-marker .lzma) capacity
 * retry.  *out is malloc'ed by the library (free with lzb_free) and holds *out_len bytes -- on error
 * the reference's partial output.  Returns LZB_RC_*; *st has the decode status. */
int lzb_decompress_alloc(lzb_ctx *ctx, int fmt, const lzb_options *opt, const uint8_t *in, size_t in_len,
                         uint8_t **out, size_t *out_len, size_t *consumed, lzb_status *st);
void lzb_free(void *p);

/* Device CRC of byte ranges (XZ block check, xz.rs:295-333): crc32[i] / crc64[i] of
 * d_data[off[i], off[i]+len[i]).  Host arrays for off/len/results. */
int lzb_crc_device(lzb_ctx *ctx, const uint8_t *d_data, const uint64_t *off, const uint64_t *len, uint32_t n,
                   uint32_t *crc32, uint64_t *crc64, void *cuda_stream);

/* ---- compress side: lzma_rs::{lzma_compress, lzma_compress_with_options, lzma2_compress, xz_compress}
 * (src/lib.rs:63-80, 91-97, 108-110).  The reference's encoders are format writers, not compressors: lzma_compress
 * emits literals only (src/encode/dumbencoder.rs), lzma2_compress / xz_compress emit stored chunks only
 * (src/encode/lzma2.rs, src/encode/xz.rs); the GPU writes the same bytes. ---- */
typedef struct lzb_compress_options { /* compress::Options / UnpackedSize, src/encode/options.rs:1-30 (LZB_FMT_LZMA only) */
    uint8_t skip_size_field;           /* UnpackedSize::SkipWritingToHeader */
    uint8_t has_value;                 /* WriteToHeader(Some(value)): no end marker; WriteToHeader(None): marker */
    uint8_t reserved[6];
    uint64_t value;
} lzb_compress_options;
/* Output capacity stream of in_len bytes needs: exact for LZB_FMT_LZMA2 / LZB_FMT_XZ; for LZB_FMT_LZMA a bound that
 * holds for any realistic input (in_len * 5/4 + 128) -- an adversarial input that needs more reports LZB_E_CAPACITY
 * with the exact size in a0. */
uint64_t lzb_encode_bound(int fmt, const lzb_compress_options *opt, uint64_t in_len);
/* Encode n independent plaintexts from HOST memory into HOST memory: stream i = in[in_off[i], in_off[i+1]) ->
 * out[out_off[i], ...), capacity out_off[i+1]-out_off[i]; out_len[i] = bytes written; st[i].code LZB_OK or
 * LZB_E_CAPACITY (a0 = bytes needed).  opt may be NULL (= Options::default()). */
int lzb_encode_batch(lzb_ctx *ctx, int fmt, const lzb_compress_options *opt, const uint8_t *in, const uint64_t *in_off,
                     uint32_t n, uint8_t *out, const uint64_t *out_off, uint64_t *out_len, lzb_status *st);
/* Same with DEVICE buffers (d_in readable up to the next multiple of 4 bytes past each plaintext); offsets and
 * results are host arrays.  Runs on `cuda_stream` (NULL = the ctx's stream) and synchronises it. */
int lzb_encode_batch_device(lzb_ctx *ctx, int fmt, const lzb_compress_options *opt, const uint8_t *d_in,
                            const uint64_t *in_off, uint32_t n, uint8_t *d_out, const uint64_t *out_off,
                            uint64_t *out_len, lzb_status *st, void *cuda_stream);

/* Renders the reference's Display string for a status ("lzma error: ...", src/error.rs:28-37) into
 * buf (NUL-terminated, truncated to buf_len).  Returns the untruncated length. */
size_t lzb_format_error(const lzb_status *st, char *buf, size_t buf_len);

#ifdef __cplusplus
}
#endif
#endif /* LZMA_B200_H */
