/*
 * lzma_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the decode path of gendx/lzma-rs @ 1f14478
 * (src/decode/{rangecoder,lzma,lzbuffer,lzma2,xz}.rs, src/xz/{mod,header,footer,crc}.rs, src/decode/util.rs,
 * src/decode/options.rs, src/error.rs).  It exists to be the bit-exact checker for the
 * CUDA path and the "lzma-rs-equivalent" CPU baseline of bench.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product library (lzma_rs_b200/csrc) never links,
 * loads or calls anything in oracle/.
 *
 * PARITY PINNING: the reference is Rust and there is no rustc/cargo in this image, so
 * the reference itself cannot be executed here.  The oracle is pinned against every
 * golden vector the reference's own tests hold for this path (tests/lzma.rs,
 * tests/xz.rs, src/decode/stream.rs tests; fixtures under tests/files) and
 * cross-checked against liblzma 5.4.5 (Python `lzma`), the same differential oracle the
 * reference's tests use (tests/lzma.rs:109-114, fuzz/fuzz_targets/compare_xz.rs:28-37).
 * Also restated: the incremental decoder decompress::Stream (lzo_stream_*, pinned on the known answers of
 * src/decode/stream.rs:348-499) and, in lzma_oracle_enc.c, the reference's three encoders (UNPINNED: the reference
 * holds no golden compressed bytes; checked by round trips and against liblzma's decoders).
 */
#ifndef LZMA_ORACLE_H
#define LZMA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error::Error variants, src/error.rs:7-17 */
enum {
    LZO_OK = 0,
    LZO_ERR_IO = 1,               /* Error::IoError        -> "io error: {}"         */
    LZO_ERR_HEADER_TOO_SHORT = 2, /* Error::HeaderTooShort -> "header too short: {}" */
    LZO_ERR_LZMA = 3,             /* Error::LzmaError      -> "lzma error: {}"       */
    LZO_ERR_XZ = 4                /* Error::XzError        -> "xz error: {}"         */
};

typedef struct {
    int kind;      /* LZO_* */
    char msg[320]; /* payload of the variant, WITHOUT the Display prefix */
} lzo_error;

/* decompress::Options, src/decode/options.rs:3-43 (allow_incomplete is stream-API only) */
typedef struct {
    int unpacked_mode;   /* 0 ReadFromHeader, 1 ReadHeaderButUseProvided(x), 2 UseProvided(x) */
    int has_provided;    /* x = Some(provided) / None */
    uint64_t provided;
    int has_memlimit;    /* memlimit: Option<usize> */
    uint64_t memlimit;
} lzo_options;

/* Result of one decode.  `out`/`out_len` = exactly the bytes the reference would have
 * handed to the caller's io::Write sink (including partial output on error).
 * `consumed` = bytes the reference would have consumed from the caller's BufRead
 * (meaningful on success). */
typedef struct {
    uint8_t *out;
    size_t out_len;
    size_t consumed;
    lzo_error err;
} lzo_result;

/* lib.rs:44-60 / 83-88 / 100-105.  opt may be NULL (= Options::default()).
 * Return value = err.kind.  Release result buffers with lzo_result_free. */
int lzo_lzma_decompress(const uint8_t *in, size_t in_len, const lzo_options *opt, lzo_result *res);
int lzo_lzma2_decompress(const uint8_t *in, size_t in_len, lzo_result *res);
int lzo_xz_decompress(const uint8_t *in, size_t in_len, lzo_result *res);
void lzo_result_free(lzo_result *res);

/* decompress::raw decoder objects: the DecoderState survives between decompress() calls (lzma.rs:597-648, lzma2.rs:11-82) */
typedef struct lzo_raw lzo_raw;
lzo_raw *lzo_raw_new(int fmt, uint32_t lc, uint32_t lp, uint32_t pb, uint32_t dict_size, int has_unpacked,
                     uint64_t unpacked, int has_memlimit, uint64_t memlimit);
void lzo_raw_reset(lzo_raw *r, int set_unpacked, int has_unpacked, uint64_t unpacked);
int lzo_raw_decompress(lzo_raw *r, const uint8_t *in, size_t in_len, lzo_result *res);
void lzo_raw_free(lzo_raw *r);

/* Formats the Display string of the error ("lzma error: ..."), src/error.rs:28-37. */
void lzo_error_display(const lzo_error *err, char *buf, size_t buf_len);

/* Multi-threaded batch driver used only for the CPU baseline timing in bench.py:
 * decodes stream i = in[in_off[i] .. in_off[i+1]) with `fmt` (0 LZMA, 1 LZMA2, 2 XZ) into
 * out[out_off[i] ..] (capacity out_off[i+1]-out_off[i]); one stream per task on
 * `nthreads` pthreads.  Writes out_len[i] and kinds[i] (LZO_*; LZO_ERR_IO with out_len
 * = needed size if the capacity was too small).  Returns number of failed streams. */
int lzo_decompress_batch(int fmt, const uint8_t *in, const uint64_t *in_off, uint32_t n,
                         uint8_t *out, const uint64_t *out_off, uint64_t *out_len,
                         int32_t *kinds, int nthreads);

/* decompress::Stream (feature `stream`, src/decode/stream.rs): the incremental push decoder, restated with its dry-run
 * logic (lzma.rs:408-524), to check the product's buffering facade.  write returns the error kind (0 = Ok, *consumed =
 * bytes accepted); finish frees the stream and reports the sink contents. */
typedef struct lzo_stream lzo_stream;
lzo_stream *lzo_stream_new(const lzo_options *opt, int allow_incomplete);
int lzo_stream_write(lzo_stream *z, const uint8_t *data, size_t n, size_t *consumed, lzo_error *err);
int lzo_stream_finish(lzo_stream *z, lzo_result *res);

/* ---- compress side (lzma_oracle_enc.c): src/lib.rs:63-80, 91-97, 108-110.  PARITY UNPINNED: the reference's tests hold
 * no golden compressed vectors, only round trips; see the header of lzma_oracle_enc.c. ---- */
/* compress::Options / compress::UnpackedSize, src/encode/options.rs:1-30 */
typedef struct {
    int skip_size_field; /* UnpackedSize::SkipWritingToHeader */
    int has_value;       /* WriteToHeader(Some(value)) : no end marker; WriteToHeader(None) : 0xFFFF.. + end marker */
    uint64_t value;
} lzo_compress_options;
int lzo_lzma_compress(const uint8_t *in, size_t n, const lzo_compress_options *opt, uint8_t **out, size_t *out_len);
int lzo_lzma2_compress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);
int lzo_xz_compress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);
void lzo_buffer_free(uint8_t *p);

/* CRC-32/ISO-HDLC and CRC-64/XZ (crate `crc` 3.x catalogue entries used by src/xz/crc.rs:3-4) */
uint32_t lzo_crc32(const uint8_t *p, size_t n);
uint64_t lzo_crc64(const uint8_t *p, size_t n);

#ifdef __cplusplus
}
#endif
#endif
