/*
 * lzma_oracle_enc.c -- CPU ORACLE, compress side (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the three encoders of gendx/lzma-rs @ 1f14478 (src/lib.rs:63-80, 91-97, 108-110).  They are
 * format writers rather than compressors: `lzma_compress` emits literals only (src/encode/dumbencoder.rs),
 * `lzma2_compress` / `xz_compress` emit stored chunks only (src/encode/lzma2.rs, src/encode/xz.rs).  Input is a byte
 * slice (what the reference's tests pass: &[u8] / Cursor), output a malloc'ed buffer.
 *
 * PARITY: pinned on the one golden COMPRESSED vector the reference holds that its own encoder produced -- the 23 bytes
 * of tests/lzma.rs:197-207 (decompress_empty_world) = src/decode/stream.rs:393,444 are exactly lzma_compress(b"")
 * (header, end marker, flush; tests/test_oracle_encoders.py::test_encoder_known_answer_empty_world) -- and on the
 * indirect known answer of src/decode/stream.rs:474-499 (half of lzma_compress(small.txt) decodes to small.txt[..26]).
 * Everything else the reference's tests do on the compress side is a round trip (tests/lzma.rs:16-28, tests/lzma2.rs,
 * tests/xz.rs:30-52); the restatement is checked the same way (its output decodes to the input with the decode oracle
 * and with liblzma, at every chunk-size edge).  No golden vector exists for literal payloads or for the stored-chunk
 * formats: for those the pin is the round trip through two independent decoders.
 */
#include <stdlib.h>
#include <string.h>

#include "lzma_oracle.h"

typedef struct {
    uint8_t *data;
    size_t len, cap;
} obuf;

static void ob_push(obuf *b, uint8_t x) {
    if (b->len == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 256;
        b->data = (uint8_t *)realloc(b->data, b->cap);
    }
    b->data[b->len++] = x;
}
static void ob_append(obuf *b, const uint8_t *p, size_t n) {
    for (size_t i = 0; i < n; i++) ob_push(b, p[i]);
}
static void ob_u32le(obuf *b, uint32_t v) {
    for (int i = 0; i < 4; i++) ob_push(b, (uint8_t)(v >> (8 * i)));
}

/* ---- RangeEncoder (src/encode/rangecoder.rs:7-106) ---- */
typedef struct {
    obuf *out;
    uint32_t range;
    uint64_t low;
    uint8_t cache;
    uint32_t cachesz;
} rangeenc;

static void re_new(rangeenc *e, obuf *out) { /* 20-32 */
    e->out = out;
    e->range = 0xFFFFFFFFu;
    e->low = 0;
    e->cache = 0;
    e->cachesz = 1;
}

static void re_write_low(rangeenc *e) { /* 34-52 */
    if (e->low < 0xFF000000ull || e->low > 0xFFFFFFFFull) {
        uint8_t tmp = e->cache;
        for (;;) {
            ob_push(e->out, (uint8_t)(tmp + (uint8_t)(e->low >> 32)));
            tmp = 0xFF;
            e->cachesz -= 1;
            if (e->cachesz == 0) break;
        }
        e->cache = (uint8_t)(e->low >> 24);
    }
    e->cachesz += 1;
    e->low = (e->low << 8) & 0xFFFFFFFFull;
}

static void re_finish(rangeenc *e) { /* 54-61 */
    for (int i = 0; i < 5; i++) re_write_low(e);
}

static void re_encode_bit(rangeenc *e, uint16_t *prob, int bit) { /* 86-106, normalize 63-84 */
    uint32_t bound = (e->range >> 11) * (uint32_t)*prob;
    if (bit) {
        *prob -= *prob >> 5;
        e->low += bound;
        e->range -= bound;
    } else {
        *prob += (uint16_t)((0x800 - *prob) >> 5);
        e->range = bound;
    }
    while (e->range < 0x01000000u) {
        e->range <<= 8;
        re_write_low(e);
    }
}

/* ---- lzma_compress_with_options (src/lib.rs:72-80 -> src/encode/dumbencoder.rs:24-139) ---- */
int lzo_lzma_compress(const uint8_t *in, size_t n, const lzo_compress_options *opt, uint8_t **out, size_t *out_len) {
    static const lzo_compress_options defaults = {0, 0, 0};
    if (!opt) opt = &defaults;
    obuf o = {0, 0, 0};
    /* from_stream, 24-62: props lc3 lp0 pb2, dict 0x800000, size field unless SkipWritingToHeader */
    ob_push(&o, (uint8_t)(3 + 9 * (0 + 5 * 2)));
    ob_u32le(&o, 0x00800000u);
    if (!opt->skip_size_field) {
        uint64_t v = opt->has_value ? opt->value : 0xFFFFFFFFFFFFFFFFull;
        for (int i = 0; i < 8; i++) ob_push(&o, (uint8_t)(v >> (8 * i)));
    }
    rangeenc e;
    re_new(&e, &o);
    static uint16_t lit[8][0x300];
    uint16_t (*literal_probs)[0x300] = (uint16_t(*)[0x300])malloc(sizeof lit);
    uint16_t is_match[4];
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 0x300; j++) literal_probs[i][j] = 0x400;
    for (int i = 0; i < 4; i++) is_match[i] = 0x400;
    /* process, 64-85 */
    uint8_t prev = 0;
    size_t input_len = 0;
    for (size_t k = 0; k < n; k++) {
        input_len = k;
        re_encode_bit(&e, &is_match[k & 3], 0);
        /* encode_literal, 124-139 */
        uint16_t *probs = literal_probs[prev >> 5];
        size_t result = 1;
        for (int i = 0; i < 8; i++) {
            int bit = (in[k] >> (7 - i)) & 1;
            re_encode_bit(&e, &probs[result], bit);
            result = (result << 1) ^ (size_t)bit;
        }
        prev = in[k];
    }
    /* finish(input_len + 1), 87-122: end marker only for WriteToHeader(None).  (input_len is the index of the last byte,
     * so the marker's pos_state is (n & 3) for n >= 1 and 1 for empty input -- as in the reference.) */
    if (!opt->skip_size_field && !opt->has_value) {
        uint16_t fresh;
        re_encode_bit(&e, &is_match[(input_len + 1) & 3], 1);
        fresh = 0x400, re_encode_bit(&e, &fresh, 0); /* new distance */
        for (int i = 0; i < 4; i++) fresh = 0x400, re_encode_bit(&e, &fresh, 0);  /* len = 0 */
        for (int i = 0; i < 6; i++) fresh = 0x400, re_encode_bit(&e, &fresh, 1);  /* pos_slot = 63 */
        for (int i = 0; i < 30; i++) fresh = 0x400, re_encode_bit(&e, &fresh, 1); /* distance 0xFFFFFFFF */
    }
    re_finish(&e);
    free(literal_probs);
    *out = o.data;
    *out_len = o.len;
    return 0;
}

/* ---- lzma2_compress (src/lib.rs:91-97 -> src/encode/lzma2.rs:4-26): stored chunks of <= 0x10000 bytes ---- */
static void lzma2_stored(obuf *o, const uint8_t *in, size_t n) {
    size_t pos = 0;
    for (;;) {
        size_t k = n - pos < 0x10000 ? n - pos : 0x10000; /* input.read(&mut buf) on a slice */
        if (k == 0) {
            ob_push(o, 0);
            break;
        }
        ob_push(o, 1);
        ob_push(o, (uint8_t)((k - 1) >> 8)); /* write_u16::<BigEndian>(n - 1) */
        ob_push(o, (uint8_t)(k - 1));
        ob_append(o, in + pos, k);
        pos += k;
    }
}

int lzo_lzma2_compress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
    obuf o = {0, 0, 0};
    lzma2_stored(&o, in, n);
    *out = o.data;
    *out_len = o.len;
    return 0;
}

/* ---- xz_compress (src/lib.rs:108-110 -> src/encode/xz.rs:9-183): one block, CheckMethod::None ---- */
static void multibyte(obuf *o, uint64_t v) { /* xz.rs:166-183 */
    for (;;) {
        uint8_t b = v & 0x7F;
        v >>= 7;
        if (v == 0) {
            ob_push(o, b);
            break;
        }
        ob_push(o, 0x80 | b);
    }
}

int lzo_xz_compress(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
    obuf o = {0, 0, 0};
    static const uint8_t magic[6] = {0xFD, 0x37, 0x7A, 0x58, 0x5A, 0x00};
    const uint8_t flags[2] = {0x00, 0x00}; /* StreamFlags { check_method: None }, xz/mod.rs:41-51 */
    /* write_header, xz.rs:31-45 */
    ob_append(&o, magic, 6);
    ob_append(&o, flags, 2);
    ob_u32le(&o, lzo_crc32(flags, 2));
    /* write_block, xz.rs:70-122 */
    size_t block_start = o.len;
    static const uint8_t bh[8] = {8 >> 2, 0x00, 0x21, 1, 22, 0, 0, 0};
    ob_append(&o, bh, 8);
    ob_u32le(&o, lzo_crc32(bh, 8));
    lzma2_stored(&o, in, n);
    size_t unpadded = o.len - block_start;
    for (size_t pad = ((unpadded ^ 3) + 1) & 3; pad; pad--) ob_push(&o, 0);
    /* write_index, xz.rs:124-164 */
    size_t index_start = o.len;
    ob_push(&o, 0);
    multibyte(&o, 1);
    multibyte(&o, unpadded);
    multibyte(&o, n);
    for (size_t pad = (((o.len - index_start) ^ 3) + 1) & 3; pad; pad--) ob_push(&o, 0);
    ob_u32le(&o, lzo_crc32(o.data + index_start, o.len - index_start));
    size_t index_size = o.len - index_start;
    /* write_footer, xz.rs:47-68 */
    uint8_t fb[6];
    uint32_t backward = (uint32_t)((index_size >> 2) - 1);
    for (int i = 0; i < 4; i++) fb[i] = (uint8_t)(backward >> (8 * i));
    fb[4] = flags[0], fb[5] = flags[1];
    ob_u32le(&o, lzo_crc32(fb, 6));
    ob_append(&o, fb, 6);
    ob_push(&o, 0x59);
    ob_push(&o, 0x5A);
    *out = o.data;
    *out_len = o.len;
    return 0;
}

void lzo_buffer_free(uint8_t *p) { free(p); }
