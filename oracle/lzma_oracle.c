/*
 * lzma_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See lzma_oracle.h.
 *
 * Plain-C restatement of the decode path of gendx/lzma-rs @ 1f14478.  Every function
 * cites the reference file:line it follows (paths relative to the reference root).
 * The reference reads from an io::BufRead and writes to an io::Write; here the reader is
 * a byte slice with a cursor and the sink is a growable byte vector, which is exactly
 * what the reference's own tests use (&[u8] / Cursor / Vec<u8>).
 *
 * Integer behaviour notes: the reference is built in release mode by its users, so
 * u32 arithmetic that could overflow (xz.rs:52 `(backward_size + 1) << 2`) wraps here.
 */
#define _GNU_SOURCE
#include "lzma_oracle.h"

#include <malloc.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * errors (src/error.rs:7-37)
 * ---------------------------------------------------------------------------------------- */
#define IO_EOF_MSG "failed to fill whole buffer" /* std::io::Read::read_exact on EOF */

static int fail(lzo_error *e, int kind, const char *fmt, ...) {
    va_list ap;
    e->kind = kind;
    va_start(ap, fmt);
    vsnprintf(e->msg, sizeof e->msg, fmt, ap);
    va_end(ap);
    return kind;
}

void lzo_error_display(const lzo_error *err, char *buf, size_t n) {
    static const char *pfx[] = {"", "io error: ", "header too short: ", "lzma error: ", "xz error: "};
    int k = err->kind;
    if (k < 0 || k > 4) k = 0;
    snprintf(buf, n, "%s%s", pfx[k], err->msg);
}

#define TRY(x)                \
    do {                      \
        int rc_ = (x);        \
        if (rc_) return rc_;  \
    } while (0)

/* ------------------------------------------------------------------------------------------
 * reader: the caller's BufRead, plus io::Take (lzma2.rs:189, xz.rs:212) as an absolute limit
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *p;
    size_t pos, end;
} reader;

/* A view = (reader, limit).  limit == r->end for the bare reader. */
typedef struct {
    reader *r;
    size_t lim;
} view;

static view view_all(reader *r) {
    view v = {r, r->end};
    return v;
}
static view view_take(reader *r, uint64_t n) { /* io::Read::take */
    view v;
    v.r = r;
    v.lim = (n > (uint64_t)(r->end - r->pos)) ? r->end : r->pos + (size_t)n;
    return v;
}
static int v_eof(const view *v) { return v->r->pos >= v->lim; } /* util.rs:9-12 is_eof */
/* byteorder read_u8 = read_exact of 1 byte; returns 0 ok / 1 UnexpectedEof */
static int v_u8(view *v, uint8_t *b) {
    if (v->r->pos >= v->lim) return 1;
    *b = v->r->p[v->r->pos++];
    return 0;
}
/* read_exact: consumes what is there even when it fails (std semantics) */
static int v_exact(view *v, uint8_t *dst, size_t n) {
    size_t avail = v->lim - v->r->pos;
    if (avail < n) {
        if (dst) memcpy(dst, v->r->p + v->r->pos, avail);
        v->r->pos += avail;
        return 1;
    }
    if (dst) memcpy(dst, v->r->p + v->r->pos, n);
    v->r->pos += n;
    return 0;
}
static int v_u16be(view *v, uint16_t *x) {
    uint8_t b[2];
    if (v_exact(v, b, 2)) return 1;
    *x = (uint16_t)((b[0] << 8) | b[1]);
    return 0;
}
static int v_u32be(view *v, uint32_t *x) {
    uint8_t b[4];
    if (v_exact(v, b, 4)) return 1;
    *x = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    return 0;
}
static int v_u32le(view *v, uint32_t *x) {
    uint8_t b[4];
    if (v_exact(v, b, 4)) return 1;
    *x = ((uint32_t)b[3] << 24) | ((uint32_t)b[2] << 16) | ((uint32_t)b[1] << 8) | b[0];
    return 0;
}
static int v_u64le(view *v, uint64_t *x) {
    uint8_t b[8];
    int i;
    if (v_exact(v, b, 8)) return 1;
    *x = 0;
    for (i = 7; i >= 0; i--) *x = (*x << 8) | b[i];
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * sink: the caller's io::Write (a Vec<u8>)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t *data;
    size_t len, cap;
} bytes;

/* Allocation hint for the batch driver (bench.py's CPU baseline): first capacity of every growable buffer of the
 * current thread, so that a many-thread run is not serialised by realloc/mmap traffic.  0 = start small. */
static __thread size_t tl_capacity_hint = 0;

static void bytes_reserve(bytes *b, size_t extra) {
    if (b->len + extra > b->cap) {
        size_t nc = b->cap ? b->cap : (tl_capacity_hint > 4096 ? tl_capacity_hint : 4096);
        while (nc < b->len + extra) nc *= 2;
        b->data = (uint8_t *)realloc(b->data, nc);
        if (!b->data) abort();
        b->cap = nc;
    }
}
static void bytes_append(bytes *b, const uint8_t *p, size_t n) {
    if (!n) return;
    bytes_reserve(b, n);
    memcpy(b->data + b->len, p, n);
    b->len += n;
}
static void bytes_push(bytes *b, uint8_t x) {
    if (b->len == b->cap) bytes_reserve(b, 1);
    b->data[b->len++] = x;
}

/* ------------------------------------------------------------------------------------------
 * LzBuffer (lzbuffer.rs:4-36): Accum (39-165) and Circular (168-321)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int circular;    /* 0 = LzAccumBuffer, 1 = LzCircularBuffer */
    bytes *sink;     /* stream: W */
    bytes buf;       /* buf: Vec<u8> */
    size_t dict_size;/* circular only */
    size_t memlimit;
    size_t cursor;   /* circular only */
    size_t len;
} lzbuf;

static void lzb_init_accum(lzbuf *b, bytes *sink, size_t memlimit) { /* lzbuffer.rs:57-64 */
    memset(b, 0, sizeof *b);
    b->sink = sink;
    b->memlimit = memlimit;
}
static void lzb_init_circ(lzbuf *b, bytes *sink, size_t dict_size, size_t memlimit) { /* 190-200 */
    memset(b, 0, sizeof *b);
    b->circular = 1;
    b->sink = sink;
    b->dict_size = dict_size;
    b->memlimit = memlimit;
}
static void lzb_drop(lzbuf *b) { free(b->buf.data); }

/* LzAccumBuffer::append_bytes, lzbuffer.rs:67-70 */
static void accum_append_bytes(lzbuf *b, const uint8_t *p, size_t n) {
    bytes_append(&b->buf, p, n);
    b->len += n;
}
/* LzAccumBuffer::reset, lzbuffer.rs:73-78 */
static void accum_reset(lzbuf *b) {
    bytes_append(b->sink, b->buf.data, b->buf.len);
    b->buf.len = 0;
    b->len = 0;
}

/* LzCircularBuffer::get, lzbuffer.rs:202-204 */
static uint8_t circ_get(const lzbuf *b, size_t idx) { return idx < b->buf.len ? b->buf.data[idx] : 0; }
/* LzCircularBuffer::set, lzbuffer.rs:206-221 */
static int circ_set(lzbuf *b, size_t idx, uint8_t v, lzo_error *e) {
    size_t new_len = idx + 1;
    if (b->buf.len < new_len) {
        if (new_len <= b->memlimit) {
            bytes_reserve(&b->buf, new_len - b->buf.len);
            memset(b->buf.data + b->buf.len, 0, new_len - b->buf.len);
            b->buf.len = new_len;
        } else {
            return fail(e, LZO_ERR_LZMA, "exceeded memory limit of %zu", b->memlimit);
        }
    }
    b->buf.data[idx] = v;
    return 0;
}

/* last_or: lzbuffer.rs:89-96 (accum), 231-238 (circular) */
static uint8_t lzb_last_or(const lzbuf *b, uint8_t lit) {
    if (!b->circular) return b->buf.len ? b->buf.data[b->buf.len - 1] : lit;
    if (b->len == 0) return lit;
    return circ_get(b, (b->dict_size + b->cursor - 1) % b->dict_size);
}

/* last_n: lzbuffer.rs:98-108 (accum), 240-256 (circular) */
static int lzb_last_n(const lzbuf *b, size_t dist, uint8_t *out, lzo_error *e) {
    if (!b->circular) {
        if (dist > b->buf.len)
            return fail(e, LZO_ERR_LZMA, "Match distance %zu is beyond output size %zu", dist, b->buf.len);
        *out = b->buf.data[b->buf.len - dist];
        return 0;
    }
    if (dist > b->dict_size)
        return fail(e, LZO_ERR_LZMA, "Match distance %zu is beyond dictionary size %zu", dist, b->dict_size);
    if (dist > b->len)
        return fail(e, LZO_ERR_LZMA, "Match distance %zu is beyond output size %zu", dist, b->len);
    *out = circ_get(b, (b->dict_size + b->cursor - dist) % b->dict_size);
    return 0;
}

/* append_literal: lzbuffer.rs:110-123 (accum), 258-270 (circular) */
static int lzb_append_literal(lzbuf *b, uint8_t lit, lzo_error *e) {
    if (!b->circular) {
        size_t new_len = b->len + 1;
        if (new_len > b->memlimit) return fail(e, LZO_ERR_LZMA, "exceeded memory limit of %zu", b->memlimit);
        bytes_push(&b->buf, lit);
        b->len = new_len;
        return 0;
    }
    TRY(circ_set(b, b->cursor, lit, e));
    b->cursor += 1;
    b->len += 1;
    if (b->cursor == b->dict_size) { /* flush the whole ring to the sink on wrap, 264-267 */
        bytes_append(b->sink, b->buf.data, b->buf.len);
        b->cursor = 0;
    }
    return 0;
}

/* append_lz: lzbuffer.rs:125-143 (accum), 272-297 (circular) */
static int lzb_append_lz(lzbuf *b, size_t len, size_t dist, lzo_error *e) {
    size_t i, off;
    if (!b->circular) {
        size_t buf_len = b->buf.len;
        if (dist > buf_len)
            return fail(e, LZO_ERR_LZMA, "LZ distance %zu is beyond output size %zu", dist, buf_len);
        off = buf_len - dist;
        bytes_reserve(&b->buf, len);
        for (i = 0; i < len; i++) { /* byte by byte: dist < len replicates with period dist */
            b->buf.data[b->buf.len++] = b->buf.data[off++];
        }
        b->len += len;
        return 0;
    }
    if (dist > b->dict_size)
        return fail(e, LZO_ERR_LZMA, "LZ distance %zu is beyond dictionary size %zu", dist, b->dict_size);
    if (dist > b->len)
        return fail(e, LZO_ERR_LZMA, "LZ distance %zu is beyond output size %zu", dist, b->len);
    off = (b->dict_size + b->cursor - dist) % b->dict_size;
    for (i = 0; i < len; i++) {
        uint8_t x = circ_get(b, off);
        TRY(lzb_append_literal(b, x, e));
        off += 1;
        if (off == b->dict_size) off = 0;
    }
    return 0;
}

/* finish: lzbuffer.rs:155-159 (accum), 309-315 (circular) */
static void lzb_finish(lzbuf *b) {
    if (!b->circular) {
        bytes_append(b->sink, b->buf.data, b->buf.len);
    } else if (b->cursor > 0) {
        bytes_append(b->sink, b->buf.data, b->cursor);
    }
}

/* ------------------------------------------------------------------------------------------
 * RangeDecoder (rangecoder.rs:7-151).  I/O failure inside the bit loop is sticky: `ioerr`
 * records the UnexpectedEof that `?` would have propagated (rangecoder.rs:64).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    view in;
    uint32_t range, code;
} rangedec;

/* RangeDecoder::new, rangecoder.rs:20-30: skip one byte (value ignored), code = BE u32 */
static int rc_new(rangedec *rc, view in) {
    uint8_t skip;
    rc->in = in;
    rc->range = 0xFFFFFFFFu;
    rc->code = 0;
    if (v_u8(&rc->in, &skip)) return 1;
    if (v_u32be(&rc->in, &rc->code)) return 1;
    return 0;
}
/* is_finished_ok, rangecoder.rs:50-52 */
static int rc_is_finished_ok(const rangedec *rc) { return rc->code == 0 && v_eof(&rc->in); }

/* normalize, rangecoder.rs:60-69 -- a single conditional shift, after every bit */
static inline int rc_normalize(rangedec *rc) {
    if (rc->range < 0x01000000u) {
        uint8_t b;
        rc->range <<= 8;
        if (v_u8(&rc->in, &b)) return 1;
        rc->code = (rc->code << 8) ^ (uint32_t)b;
    }
    return 0;
}
/* get_bit / get, rangecoder.rs:72-90 */
static inline int rc_get_bit(rangedec *rc, uint32_t *bit) {
    rc->range >>= 1;
    *bit = rc->code >= rc->range;
    if (*bit) rc->code -= rc->range;
    return rc_normalize(rc);
}
static int rc_get(rangedec *rc, size_t count, uint32_t *out) {
    uint32_t result = 0, bit;
    size_t i;
    for (i = 0; i < count; i++) {
        if (rc_get_bit(rc, &bit)) return 1;
        result = (result << 1) ^ bit;
    }
    *out = result;
    return 0;
}
/* decode_bit, rangecoder.rs:93-120 (update is always true off the stream API, lzma.rs:395-401) */
static inline int rc_decode_bit(rangedec *rc, uint16_t *prob, uint32_t *bit) {
    uint32_t bound = (rc->range >> 11) * (uint32_t)*prob;
    if (rc->code < bound) {
        *prob = (uint16_t)(*prob + ((0x800u - *prob) >> 5));
        rc->range = bound;
        *bit = 0;
    } else {
        *prob = (uint16_t)(*prob - (*prob >> 5));
        rc->code -= bound;
        rc->range -= bound;
        *bit = 1;
    }
    return rc_normalize(rc);
}
/* parse_bit_tree, rangecoder.rs:122-134 */
static int rc_bit_tree(rangedec *rc, size_t num_bits, uint16_t *probs, uint32_t *out) {
    uint32_t tmp = 1, bit;
    size_t i;
    for (i = 0; i < num_bits; i++) {
        if (rc_decode_bit(rc, &probs[tmp], &bit)) return 1;
        tmp = (tmp << 1) ^ bit;
    }
    *out = tmp - (1u << num_bits);
    return 0;
}
/* parse_reverse_bit_tree, rangecoder.rs:136-151 */
static int rc_rev_bit_tree(rangedec *rc, size_t num_bits, uint16_t *probs, size_t offset, uint32_t *out) {
    uint32_t result = 0, bit;
    size_t tmp = 1, i;
    for (i = 0; i < num_bits; i++) {
        if (rc_decode_bit(rc, &probs[offset + tmp], &bit)) return 1;
        tmp = (tmp << 1) ^ bit;
        result ^= bit << i;
    }
    *out = result;
    return 0;
}

/* LenDecoder, rangecoder.rs:202-270 */
typedef struct {
    uint16_t choice, choice2;
    uint16_t low[16][8];
    uint16_t mid[16][8];
    uint16_t high[256];
} lendec;

static void fill16(uint16_t *p, size_t n) {
    size_t i;
    for (i = 0; i < n; i++) p[i] = 0x400;
}
static void lendec_init(lendec *d) { fill16((uint16_t *)d, sizeof *d / 2); }
/* LenDecoder::decode, rangecoder.rs:256-269 */
static int lendec_decode(lendec *d, rangedec *rc, size_t pos_state, size_t *len) {
    uint32_t bit, v;
    if (rc_decode_bit(rc, &d->choice, &bit)) return 1;
    if (!bit) {
        if (rc_bit_tree(rc, 3, d->low[pos_state], &v)) return 1;
        *len = v;
        return 0;
    }
    if (rc_decode_bit(rc, &d->choice2, &bit)) return 1;
    if (!bit) {
        if (rc_bit_tree(rc, 3, d->mid[pos_state], &v)) return 1;
        *len = (size_t)v + 8;
        return 0;
    }
    if (rc_bit_tree(rc, 8, d->high, &v)) return 1;
    *len = (size_t)v + 16;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * DecoderState (lzma.rs:165-593)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t lc, lp, pb; /* LzmaProperties, lzma.rs:42-59 */
    int has_unpacked;    /* unpacked_size: Option<u64> */
    uint64_t unpacked_size;
    uint16_t *literal_probs; /* Vec2D (1 << (lc+lp)) x 0x300 */
    size_t literal_rows;
    uint16_t pos_slot[4][64];
    uint16_t align[16];
    uint16_t pos_decoders[115];
    uint16_t is_match[192];
    uint16_t is_rep[12], is_rep_g0[12], is_rep_g1[12], is_rep_g2[12];
    uint16_t is_rep_0long[192];
    size_t state;
    size_t rep[4];
    lendec len_decoder, rep_len_decoder;
} decstate;

static void ds_fill_small(decstate *s) {
    fill16(&s->pos_slot[0][0], 4 * 64);
    fill16(s->align, 16);
    fill16(s->pos_decoders, 115);
    fill16(s->is_match, 192);
    fill16(s->is_rep, 12);
    fill16(s->is_rep_g0, 12);
    fill16(s->is_rep_g1, 12);
    fill16(s->is_rep_g2, 12);
    fill16(s->is_rep_0long, 192);
    s->state = 0;
    s->rep[0] = s->rep[1] = s->rep[2] = s->rep[3] = 0;
    lendec_init(&s->len_decoder);
    lendec_init(&s->rep_len_decoder);
}
/* DecoderState::new, lzma.rs:188-214 */
static void ds_new(decstate *s, uint32_t lc, uint32_t lp, uint32_t pb, int has_unpacked, uint64_t unpacked) {
    memset(s, 0, sizeof *s);
    s->lc = lc;
    s->lp = lp;
    s->pb = pb;
    s->has_unpacked = has_unpacked;
    s->unpacked_size = unpacked;
    s->literal_rows = (size_t)1 << (lc + lp);
    s->literal_probs = (uint16_t *)malloc(s->literal_rows * 0x300 * sizeof(uint16_t));
    if (!s->literal_probs) abort();
    fill16(s->literal_probs, s->literal_rows * 0x300);
    ds_fill_small(s);
}
/* DecoderState::reset_state, lzma.rs:216-249 */
static void ds_reset_state(decstate *s, uint32_t lc, uint32_t lp, uint32_t pb) {
    if (s->lc + s->lp != lc + lp) {
        free(s->literal_probs);
        s->literal_rows = (size_t)1 << (lc + lp);
        s->literal_probs = (uint16_t *)malloc(s->literal_rows * 0x300 * sizeof(uint16_t));
        if (!s->literal_probs) abort();
    }
    fill16(s->literal_probs, s->literal_rows * 0x300);
    s->lc = lc;
    s->lp = lp;
    s->pb = pb;
    ds_fill_small(s);
}
static void ds_drop(decstate *s) { free(s->literal_probs); }

#define IOFAIL(e) fail((e), LZO_ERR_IO, IO_EOF_MSG)

/* decode_literal, lzma.rs:526-561 */
static int ds_decode_literal(decstate *s, lzbuf *out, rangedec *rc, uint8_t *byte, lzo_error *e) {
    size_t prev_byte = lzb_last_or(out, 0);
    size_t result = 1;
    size_t lit_state = ((out->len & (((size_t)1 << s->lp) - 1)) << s->lc) + (prev_byte >> (8 - s->lc));
    uint16_t *probs = s->literal_probs + lit_state * 0x300;
    uint32_t bit;

    if (s->state >= 7) {
        uint8_t mb = 0;
        size_t match_byte;
        TRY(lzb_last_n(out, s->rep[0] + 1, &mb, e));
        match_byte = mb;
        while (result < 0x100) {
            size_t match_bit = (match_byte >> 7) & 1;
            match_byte <<= 1;
            if (rc_decode_bit(rc, &probs[((1 + match_bit) << 8) + result], &bit)) return IOFAIL(e);
            result = (result << 1) ^ bit;
            if (match_bit != bit) break;
        }
    }
    while (result < 0x100) {
        if (rc_decode_bit(rc, &probs[result], &bit)) return IOFAIL(e);
        result = (result << 1) ^ bit;
    }
    *byte = (uint8_t)(result - 0x100);
    return 0;
}

/* decode_distance, lzma.rs:563-592 */
static int ds_decode_distance(decstate *s, rangedec *rc, size_t length, size_t *dist, lzo_error *e) {
    size_t len_state = length > 3 ? 3 : length;
    uint32_t pos_slot, v;
    size_t num_direct_bits, result;

    if (rc_bit_tree(rc, 6, s->pos_slot[len_state], &pos_slot)) return IOFAIL(e);
    if (pos_slot < 4) {
        *dist = pos_slot;
        return 0;
    }
    num_direct_bits = (pos_slot >> 1) - 1;
    result = (size_t)(2 ^ (pos_slot & 1)) << num_direct_bits;
    if (pos_slot < 14) {
        if (rc_rev_bit_tree(rc, num_direct_bits, s->pos_decoders, result - pos_slot, &v)) return IOFAIL(e);
        result += v;
    } else {
        if (rc_get(rc, num_direct_bits - 4, &v)) return IOFAIL(e);
        result += (size_t)v << 4;
        if (rc_rev_bit_tree(rc, 4, s->align, 0, &v)) return IOFAIL(e);
        result += v;
    }
    *dist = result;
    return 0;
}

enum { ST_CONTINUE = 0, ST_FINISHED = 1 };

/* process_next_inner, lzma.rs:278-393.  update = 0 is the stream API's dry run (try_process_next, lzma.rs:408-419): the
 * decisions are computed but neither the window nor the state may change.  Within one symbol every probability is
 * used at most once, so the decisions do not depend on the updates; the caller runs the dry run on a COPY of the
 * DecoderState, and here update = 0 only has to skip what touches the window (append_literal / append_lz and the
 * errors they can raise, which the reference's dry run cannot see either). */
static int ds_process_next_ex(decstate *s, lzbuf *out, rangedec *rc, int *status, int update, lzo_error *e);
static int ds_process_next(decstate *s, lzbuf *out, rangedec *rc, int *status, lzo_error *e) {
    return ds_process_next_ex(s, out, rc, status, 1, e);
}
static int ds_process_next_ex(decstate *s, lzbuf *out, rangedec *rc, int *status, int update, lzo_error *e) {
    size_t pos_state = out->len & (((size_t)1 << s->pb) - 1);
    uint32_t bit;
    size_t len;

    *status = ST_CONTINUE;
    if (rc_decode_bit(rc, &s->is_match[(s->state << 4) + pos_state], &bit)) return IOFAIL(e);
    if (!bit) { /* literal, 287-307 */
        uint8_t byte = 0;
        TRY(ds_decode_literal(s, out, rc, &byte, e));
        if (!update) return 0;
        TRY(lzb_append_literal(out, byte, e));
        s->state = s->state < 4 ? 0 : (s->state < 10 ? s->state - 3 : s->state - 6);
        return 0;
    }

    if (rc_decode_bit(rc, &s->is_rep[s->state], &bit)) return IOFAIL(e);
    if (bit) { /* rep, 312-353 */
        if (rc_decode_bit(rc, &s->is_rep_g0[s->state], &bit)) return IOFAIL(e);
        if (!bit) {
            if (rc_decode_bit(rc, &s->is_rep_0long[(s->state << 4) + pos_state], &bit)) return IOFAIL(e);
            if (!bit) { /* short rep, 321-327 */
                s->state = s->state < 7 ? 9 : 11;
                return update ? lzb_append_lz(out, 1, s->rep[0] + 1, e) : 0;
            }
        } else {
            size_t idx, d, i;
            if (rc_decode_bit(rc, &s->is_rep_g1[s->state], &bit)) return IOFAIL(e);
            if (!bit) {
                idx = 1;
            } else {
                if (rc_decode_bit(rc, &s->is_rep_g2[s->state], &bit)) return IOFAIL(e);
                idx = bit ? 3 : 2;
            }
            d = s->rep[idx]; /* LRU rotate, 338-345 */
            for (i = idx; i > 0; i--) s->rep[i] = s->rep[i - 1];
            s->rep[0] = d;
        }
        if (lendec_decode(&s->rep_len_decoder, rc, pos_state, &len)) return IOFAIL(e);
        s->state = s->state < 7 ? 8 : 11;
    } else { /* new match, 355-383 */
        size_t rep0 = 0;
        s->rep[3] = s->rep[2];
        s->rep[2] = s->rep[1];
        s->rep[1] = s->rep[0];
        if (lendec_decode(&s->len_decoder, rc, pos_state, &len)) return IOFAIL(e);
        s->state = s->state < 7 ? 7 : 10;
        TRY(ds_decode_distance(s, rc, len, &rep0, e));
        s->rep[0] = rep0;
        if (s->rep[0] == 0xFFFFFFFFu) { /* end-of-stream marker, 373-381 */
            if (rc_is_finished_ok(rc)) {
                *status = ST_FINISHED;
                return 0;
            }
            return fail(e, LZO_ERR_LZMA, "Found end-of-stream marker but more bytes are available");
        }
    }
    len += 2;
    return update ? lzb_append_lz(out, len, s->rep[0] + 1, e) : 0;
}

/* process -> process_mode(Finish), lzma.rs:255-261, 435-455, 496-523 */
static int ds_process(decstate *s, lzbuf *out, rangedec *rc, lzo_error *e) {
    for (;;) {
        int status;
        if (s->has_unpacked) {
            if ((uint64_t)out->len >= s->unpacked_size) break;
        } else if (rc_is_finished_ok(rc)) {
            break;
        }
        TRY(ds_process_next(s, out, rc, &status, e));
        if (status == ST_FINISHED) break;
    }
    if (s->has_unpacked && s->unpacked_size != (uint64_t)out->len)
        return fail(e, LZO_ERR_LZMA, "Expected unpacked size of %llu but decompressed to %zu",
                    (unsigned long long)s->unpacked_size, out->len);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * .lzma: LzmaParams::read_header (lzma.rs:96-161), LzmaDecoder::{new,decompress} (607-648),
 * lib.rs:44-60
 * ---------------------------------------------------------------------------------------- */
static int lzma_decompress_impl(reader *r, const lzo_options *opt, bytes *sink, lzo_error *e) {
    static const lzo_options defaults = {0, 0, 0, 0, 0};
    view in = view_all(r);
    uint8_t props;
    uint32_t pb, lc, lp, dict_size;
    int has_unpacked = 0;
    uint64_t unpacked = 0;
    size_t memlimit;
    decstate st;
    lzbuf out;
    rangedec rc;
    int rcode;

    if (!opt) opt = &defaults;
    if (v_u8(&in, &props)) return fail(e, LZO_ERR_HEADER_TOO_SHORT, IO_EOF_MSG);
    pb = props;
    if (pb >= 225) return fail(e, LZO_ERR_LZMA, "LZMA header invalid properties: %u must be < 225", pb);
    lc = pb % 9;
    pb /= 9;
    lp = pb % 5;
    pb /= 5;
    if (v_u32le(&in, &dict_size)) return fail(e, LZO_ERR_HEADER_TOO_SHORT, IO_EOF_MSG);
    if (dict_size < 0x1000) dict_size = 0x1000;
    switch (opt->unpacked_mode) {
    case 0: { /* ReadFromHeader */
        uint64_t v;
        if (v_u64le(&in, &v)) return fail(e, LZO_ERR_HEADER_TOO_SHORT, IO_EOF_MSG);
        if (v != 0xFFFFFFFFFFFFFFFFull) {
            has_unpacked = 1;
            unpacked = v;
        }
        break;
    }
    case 1: { /* ReadHeaderButUseProvided(x) */
        uint64_t v;
        if (v_u64le(&in, &v)) return fail(e, LZO_ERR_HEADER_TOO_SHORT, IO_EOF_MSG);
        has_unpacked = opt->has_provided;
        unpacked = opt->provided;
        break;
    }
    default: /* UseProvided(x) */
        has_unpacked = opt->has_provided;
        unpacked = opt->provided;
        break;
    }
    memlimit = opt->has_memlimit ? (size_t)opt->memlimit : (size_t)-1; /* lzma.rs:610 */

    ds_new(&st, lc, lp, pb, has_unpacked, unpacked);
    lzb_init_circ(&out, sink, dict_size, memlimit);
    if (rc_new(&rc, in)) { /* lzma.rs:643-644 */
        rcode = fail(e, LZO_ERR_LZMA, "LZMA stream too short: " IO_EOF_MSG);
    } else {
        rcode = ds_process(&st, &out, &rc, e);
        if (!rcode) lzb_finish(&out); /* lzma.rs:646 */
    }
    lzb_drop(&out);
    ds_drop(&st);
    return rcode;
}

/* ------------------------------------------------------------------------------------------
 * decompress::Stream (feature `stream`): src/decode/stream.rs:66-346 over process_mode(Partial),
 * lzma.rs:421-524.  Used to check the buffering facade of the product (lzma_rs_b200.Stream).
 * ---------------------------------------------------------------------------------------- */
#define MAX_REQUIRED_INPUT 20 /* lzma.rs:13 */
#define MAX_TMP_LEN 18        /* stream.rs:11-24: 13 header bytes + 5 range-coder start bytes */

struct lzo_stream {
    lzo_options opt;
    int allow_incomplete;
    int state; /* 0 Header, 1 Data, 2 gone (a write failed: stream.rs:310) */
    uint8_t tmp[MAX_TMP_LEN];
    size_t tmp_len;
    bytes sink;
    decstate st;
    lzbuf out;
    uint32_t range, code;
    uint8_t partial[MAX_REQUIRED_INPUT]; /* DecoderState::partial_input_buf, lzma.rs:183-184 */
    size_t partial_len;
};

/* try_process_next, lzma.rs:408-419: a dry run of one symbol over `buf`; returns nonzero if it fails in any way */
static int ds_try_process_next(decstate *s, lzbuf *out, const uint8_t *buf, size_t n, uint32_t range, uint32_t code) {
    decstate copy = *s;
    reader r = {buf, 0, n};
    rangedec rc;
    lzo_error e;
    int status, rcode;
    size_t bytes_lit = s->literal_rows * 0x300 * sizeof(uint16_t);
    copy.literal_probs = (uint16_t *)malloc(bytes_lit);
    if (!copy.literal_probs) abort();
    memcpy(copy.literal_probs, s->literal_probs, bytes_lit);
    rc.in = view_all(&r);
    rc.range = range;
    rc.code = code;
    rcode = ds_process_next_ex(&copy, out, &rc, &status, 0, &e);
    free(copy.literal_probs);
    return rcode;
}

/* read_partial_input_buf, lzma.rs:421-433: fill the 20-byte buffer from the stream */
static void ds_read_partial(lzo_stream *z, rangedec *rc) {
    size_t avail = rc->in.lim - rc->in.r->pos, room = MAX_REQUIRED_INPUT - z->partial_len;
    size_t k = avail < room ? avail : room;
    memcpy(z->partial + z->partial_len, rc->in.r->p + rc->in.r->pos, k);
    rc->in.r->pos += k;
    z->partial_len += k;
}

/* process_mode, lzma.rs:435-524, both modes (partial = 1: ProcessingMode::Partial) */
static int ds_process_mode(lzo_stream *z, rangedec *rc, int partial, lzo_error *e) {
    decstate *s = &z->st;
    lzbuf *out = &z->out;
    for (;;) {
        int status;
        if (s->has_unpacked) {
            if ((uint64_t)out->len >= s->unpacked_size) break;
        } else if ((partial ? v_eof(&rc->in) : rc_is_finished_ok(rc)) && z->partial_len == 0) {
            break;
        }
        if (z->partial_len > 0) {
            uint8_t tmp[MAX_REQUIRED_INPUT];
            reader tr;
            rangedec trc;
            size_t used;
            ds_read_partial(z, rc);
            memcpy(tmp, z->partial, sizeof tmp);
            if (partial && z->partial_len < MAX_REQUIRED_INPUT &&
                ds_try_process_next(s, out, tmp, z->partial_len, rc->range, rc->code))
                return 0; /* need more data */
            tr.p = tmp;
            tr.pos = 0;
            tr.end = z->partial_len;
            trc.in = view_all(&tr);
            trc.range = rc->range;
            trc.code = rc->code;
            TRY(ds_process_next(s, out, &trc, &status, e));
            rc->range = trc.range;
            rc->code = trc.code;
            used = tr.pos;
            memmove(z->partial, tmp + used, z->partial_len - used);
            z->partial_len -= used;
            if (status == ST_FINISHED) break;
        } else {
            const uint8_t *buf = rc->in.r->p + rc->in.r->pos; /* fill_buf(): the rest of this write's bytes */
            size_t n = rc->in.lim - rc->in.r->pos;
            if (partial && n < MAX_REQUIRED_INPUT && ds_try_process_next(s, out, buf, n, rc->range, rc->code)) {
                ds_read_partial(z, rc);
                return 0;
            }
            TRY(ds_process_next(s, out, rc, &status, e));
            if (status == ST_FINISHED) break;
        }
    }
    if (s->has_unpacked && !partial && s->unpacked_size != (uint64_t)out->len)
        return fail(e, LZO_ERR_LZMA, "Expected unpacked size of %llu but decompressed to %zu",
                    (unsigned long long)s->unpacked_size, out->len);
    return 0;
}

lzo_stream *lzo_stream_new(const lzo_options *opt, int allow_incomplete) {
    static const lzo_options defaults = {0, 0, 0, 0, 0};
    lzo_stream *z = (lzo_stream *)calloc(1, sizeof *z);
    if (!z) abort();
    z->opt = opt ? *opt : defaults;
    z->allow_incomplete = allow_incomplete;
    return z;
}

/* Stream::read_header, stream.rs:157-190.  Returns 0 and *to_data = 1 (Data) / 0 (stay in Header: need more bytes),
 * or the fatal error kind.  Consumes from `in` exactly what the reference's reads consume. */
static int stream_read_header(lzo_stream *z, view *in, int *to_data, lzo_error *e) {
    uint8_t props;
    uint32_t pb, lc, lp, dict_size;
    int has_unpacked = 0;
    uint64_t unpacked = 0, v;
    rangedec rc;
    *to_data = 0;
    /* LzmaParams::read_header, lzma.rs:96-161: HeaderTooShort means "try again later" here */
    if (v_u8(in, &props)) return 0;
    pb = props;
    if (pb >= 225) return fail(e, LZO_ERR_LZMA, "LZMA header invalid properties: %u must be < 225", pb);
    lc = pb % 9;
    pb /= 9;
    lp = pb % 5;
    pb /= 5;
    if (v_u32le(in, &dict_size)) return 0;
    if (dict_size < 0x1000) dict_size = 0x1000;
    switch (z->opt.unpacked_mode) {
    case 0:
        if (v_u64le(in, &v)) return 0;
        if (v != 0xFFFFFFFFFFFFFFFFull) has_unpacked = 1, unpacked = v;
        break;
    case 1:
        if (v_u64le(in, &v)) return 0;
        has_unpacked = z->opt.has_provided;
        unpacked = z->opt.provided;
        break;
    default:
        has_unpacked = z->opt.has_provided;
        unpacked = z->opt.provided;
        break;
    }
    if (rc_new(&rc, *in)) return 0; /* "Failed to create a RangeDecoder because we need more data" */
    ds_new(&z->st, lc, lp, pb, has_unpacked, unpacked);
    lzb_init_circ(&z->out, &z->sink, dict_size, z->opt.has_memlimit ? (size_t)z->opt.memlimit : (size_t)-1);
    z->range = rc.range;
    z->code = rc.code;
    *to_data = 1;
    return 0;
}

/* Stream::read_data, stream.rs:192-207 */
static int stream_read_data(lzo_stream *z, reader *r, lzo_error *e) {
    rangedec rc;
    rc.in = view_all(r);
    rc.range = z->range;
    rc.code = z->code;
    TRY(ds_process_mode(z, &rc, 1, e));
    z->range = rc.range;
    z->code = rc.code;
    return 0;
}

static void stream_drop_data(lzo_stream *z) {
    lzb_drop(&z->out);
    ds_drop(&z->st);
}

/* <Stream as io::Write>::write, stream.rs:227-325.  Returns the error kind (0 = Ok(*consumed)). */
int lzo_stream_write(lzo_stream *z, const uint8_t *data, size_t n, size_t *consumed, lzo_error *e) {
    reader input = {data, 0, n};
    int rcode = 0, to_data = 0;
    e->kind = 0;
    e->msg[0] = 0;
    if (z->state == 0) {
        if (z->tmp_len > 0) { /* fill the tmp buffer, then try the header from it */
            size_t k = n < MAX_TMP_LEN - z->tmp_len ? n : MAX_TMP_LEN - z->tmp_len;
            reader tr;
            view tv;
            memcpy(z->tmp + z->tmp_len, data, k);
            input.pos = k;
            z->tmp_len += k;
            tr.p = z->tmp;
            tr.pos = 0;
            tr.end = z->tmp_len;
            tv = view_all(&tr);
            rcode = stream_read_header(z, &tv, &to_data, e);
            if (!rcode && to_data) { /* keep what follows the header + start bytes for the next call */
                memmove(z->tmp, z->tmp + tr.pos, z->tmp_len - tr.pos);
                z->tmp_len -= tr.pos;
            }
        } else {
            view iv = view_all(&input);
            rcode = stream_read_header(z, &iv, &to_data, e);
            if (!rcode && !to_data) { /* not enough bytes: remember them (partial reads are undone first) */
                size_t k = n < MAX_TMP_LEN ? n : MAX_TMP_LEN;
                input.pos = 0;
                memcpy(z->tmp, data, k);
                input.pos = k;
                z->tmp_len = k;
            }
        }
        if (rcode) {
            z->state = 2;
            return rcode;
        }
        if (to_data) z->state = 1;
    } else if (z->state == 1) {
        if (z->tmp_len > 0) {
            reader tr = {z->tmp, 0, z->tmp_len};
            rcode = stream_read_data(z, &tr, e);
            z->tmp_len = 0;
        }
        if (!rcode) rcode = stream_read_data(z, &input, e);
        if (rcode) {
            stream_drop_data(z);
            z->state = 2;
            return rcode;
        }
    }
    *consumed = input.pos;
    return 0;
}

/* Stream::finish, stream.rs:119-151.  Frees the stream.  res->out = what reached the sink (the reference returns the
 * sink only on success; on error it is reported here as well, for comparison with the facade's partial output). */
int lzo_stream_finish(lzo_stream *z, lzo_result *res) {
    int rcode = 0;
    memset(res, 0, sizeof *res);
    if (z->state == 2) {
        rcode = fail(&res->err, LZO_ERR_LZMA, "can't finish stream because of previous write error");
    } else if (z->state == 0) {
        if (z->tmp_len > 0) rcode = fail(&res->err, LZO_ERR_LZMA, "failed to read header");
    } else {
        if (!z->allow_incomplete) {
            reader tr = {z->tmp, 0, z->tmp_len};
            rangedec rc;
            rc.in = view_all(&tr);
            rc.range = z->range;
            rc.code = z->code;
            rcode = ds_process_mode(z, &rc, 0, &res->err);
        }
        if (!rcode) lzb_finish(&z->out);
        stream_drop_data(z);
    }
    res->out = z->sink.data;
    res->out_len = z->sink.len;
    free(z);
    return rcode;
}

/* ------------------------------------------------------------------------------------------
 * LZMA2: Lzma2Decoder (lzma2.rs:11-229)
 * ---------------------------------------------------------------------------------------- */
/* parse_uncompressed, lzma2.rs:195-229 */
static int lzma2_parse_uncompressed(lzbuf *accum, view *in, int reset_dict, lzo_error *e) {
    uint16_t us;
    size_t unpacked_size, avail;
    if (v_u16be(in, &us)) return fail(e, LZO_ERR_LZMA, "LZMA2 expected unpacked size: " IO_EOF_MSG);
    unpacked_size = (size_t)us + 1;
    if (reset_dict) accum_reset(accum);
    avail = in->lim - in->r->pos;
    if (avail < unpacked_size) {
        in->r->pos += avail;
        return fail(e, LZO_ERR_LZMA, "LZMA2 expected %zu uncompressed bytes: " IO_EOF_MSG, unpacked_size);
    }
    accum_append_bytes(accum, in->r->p + in->r->pos, unpacked_size);
    in->r->pos += unpacked_size;
    return 0;
}

/* parse_lzma, lzma2.rs:84-193 */
static int lzma2_parse_lzma(decstate *st, lzbuf *accum, view *in, uint8_t status, lzo_error *e) {
    int reset_dict, reset_state, reset_props;
    uint16_t u16v;
    uint64_t unpacked_size, packed_size;
    rangedec rc;
    view taken;

    if ((status & 0x80) == 0)
        return fail(e, LZO_ERR_LZMA, "LZMA2 invalid status %u, must be 0, 1, 2 or >= 128", (unsigned)status);
    switch ((status >> 5) & 3) {
    case 0: reset_dict = 0; reset_state = 0; reset_props = 0; break;
    case 1: reset_dict = 0; reset_state = 1; reset_props = 0; break;
    case 2: reset_dict = 0; reset_state = 1; reset_props = 1; break;
    default: reset_dict = 1; reset_state = 1; reset_props = 1; break;
    }
    if (v_u16be(in, &u16v)) return fail(e, LZO_ERR_LZMA, "LZMA2 expected unpacked size: " IO_EOF_MSG);
    unpacked_size = ((((uint64_t)(status & 0x1F)) << 16) | (uint64_t)u16v) + 1;
    if (v_u16be(in, &u16v)) return fail(e, LZO_ERR_LZMA, "LZMA2 expected packed size: " IO_EOF_MSG);
    packed_size = (uint64_t)u16v + 1;

    if (reset_dict) accum_reset(accum);
    if (reset_state) {
        uint32_t lc = st->lc, lp = st->lp, pb = st->pb;
        if (reset_props) {
            uint8_t props;
            if (v_u8(in, &props)) return fail(e, LZO_ERR_LZMA, "LZMA2 expected new properties: " IO_EOF_MSG);
            pb = props;
            if (pb >= 225) return fail(e, LZO_ERR_LZMA, "LZMA2 invalid properties: %u must be < 225", pb);
            lc = pb % 9;
            pb /= 9;
            lp = pb % 5;
            pb /= 5;
            if (lc + lp > 4)
                return fail(e, LZO_ERR_LZMA, "LZMA2 invalid properties: lc + lp (%u + %u) must be <= 4", lc, lp);
        }
        ds_reset_state(st, lc, lp, pb);
    }
    st->has_unpacked = 1; /* set_unpacked_size(Some(unpacked_size + accum.len())), 186-187 */
    st->unpacked_size = unpacked_size + (uint64_t)accum->len;

    taken = view_take(in->r, packed_size); /* 189 */
    if (taken.lim > in->lim) taken.lim = in->lim;
    if (rc_new(&rc, taken)) return fail(e, LZO_ERR_LZMA, "LZMA input too short: " IO_EOF_MSG);
    /* NB: the Take is dropped without draining (lzma2.rs:189-192): whatever the range decoder
     * left unread stays in `in` and is parsed as the next control byte. */
    return ds_process(st, accum, &rc, e);
}

/* Lzma2Decoder::decompress, lzma2.rs:52-82, on a DecoderState that outlives the call (the raw decoder object) */
static int lzma2_decompress_state(view *in, bytes *sink, decstate *stp, lzo_error *e);

/* Lzma2Decoder::new + decompress, lzma2.rs:23-34, 52-82 */
static int lzma2_decompress_view(view *in, bytes *sink, lzo_error *e) {
    decstate st;
    int rcode;
    ds_new(&st, 0, 0, 0, 0, 0);
    rcode = lzma2_decompress_state(in, sink, &st, e);
    ds_drop(&st);
    return rcode;
}

static int lzma2_decompress_state(view *in, bytes *sink, decstate *stp, lzo_error *e) {
    lzbuf accum;
    int rcode = 0;
#define st (*stp)
    lzb_init_accum(&accum, sink, (size_t)-1);
    for (;;) {
        uint8_t status;
        if (v_u8(in, &status)) {
            rcode = fail(e, LZO_ERR_LZMA, "LZMA2 expected new status: " IO_EOF_MSG);
            break;
        }
        if (status == 0) break;
        if (status == 1)
            rcode = lzma2_parse_uncompressed(&accum, in, 1, e);
        else if (status == 2)
            rcode = lzma2_parse_uncompressed(&accum, in, 0, e);
        else
            rcode = lzma2_parse_lzma(&st, &accum, in, status, e);
        if (rcode) break;
    }
    if (!rcode) lzb_finish(&accum); /* lzma2.rs:80 */
    lzb_drop(&accum);
#undef st
    return rcode;
}

/* ------------------------------------------------------------------------------------------
 * CRC-32/ISO-HDLC and CRC-64/XZ -- crate `crc` 3.x catalogue algorithms named at
 * src/xz/crc.rs:3-4 (reflected; init = xorout = all ones; poly 0x04C11DB7 / 0x42F0E1EBA9EA3693).
 * Check values: crc32("123456789") = 0xCBF43926, crc64 = 0x995DC9BBDF1939FA.
 * ---------------------------------------------------------------------------------------- */
static uint32_t crc32_tab[256];
static uint64_t crc64_tab[256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;
static void crc_init(void) {
    uint32_t i;
    int k;
    for (i = 0; i < 256; i++) {
        uint32_t c = i;
        uint64_t d = i;
        for (k = 0; k < 8; k++) {
            c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
            d = (d & 1) ? (d >> 1) ^ 0xC96C5795D7870F42ull : d >> 1;
        }
        crc32_tab[i] = c;
        crc64_tab[i] = d;
    }
}
static uint32_t crc32_update(uint32_t c, const uint8_t *p, size_t n) { /* c is the pre-inverted register */
    while (n--) c = crc32_tab[(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c;
}
uint32_t lzo_crc32(const uint8_t *p, size_t n) {
    pthread_once(&crc_once, crc_init);
    return crc32_update(0xFFFFFFFFu, p, n) ^ 0xFFFFFFFFu;
}
uint64_t lzo_crc64(const uint8_t *p, size_t n) {
    uint64_t c = ~0ull;
    pthread_once(&crc_once, crc_init);
    while (n--) c = crc64_tab[(c ^ *p++) & 0xFF] ^ (c >> 8);
    return ~c;
}

/* ------------------------------------------------------------------------------------------
 * XZ container (xz.rs:18-464, xz/header.rs:20-51, xz/mod.rs:18-76, xz/footer.rs:4)
 * ---------------------------------------------------------------------------------------- */
enum { CHECK_NONE = 0x00, CHECK_CRC32 = 0x01, CHECK_CRC64 = 0x04, CHECK_SHA256 = 0x0A };

static const char *check_name(int m) { /* #[derive(Debug)] on CheckMethod, xz/mod.rs:53-60 */
    switch (m) {
    case CHECK_NONE: return "None";
    case CHECK_CRC32: return "Crc32";
    case CHECK_CRC64: return "Crc64";
    default: return "Sha256";
    }
}

/* StreamFlags::parse, xz/mod.rs:24-39 + CheckMethod::try_from 62-76 */
static int stream_flags_parse(uint16_t field, int *check, lzo_error *e) {
    uint8_t b0 = (uint8_t)(field >> 8), b1 = (uint8_t)field;
    if (b0 != 0) return fail(e, LZO_ERR_XZ, "Invalid null byte in Stream Flags: %x", (unsigned)b0);
    if (b1 != CHECK_NONE && b1 != CHECK_CRC32 && b1 != CHECK_CRC64 && b1 != CHECK_SHA256)
        return fail(e, LZO_ERR_XZ, "Invalid check method %x, expected one of [0x00, 0x01, 0x04, 0x0A]", (unsigned)b1);
    *check = b1;
    return 0;
}

/* A reader over a byte window that digests (CRC32) every byte it hands out -- models
 * BufReader<CrcDigestRead<Take<..>>> at xz.rs:211-214: the BufReader pulls the whole Take in
 * one read, so on success every header byte is digested; `crc` is the running register. */
typedef struct {
    const uint8_t *p;
    size_t pos, end;
} hdr_reader;
static int h_u8(hdr_reader *h, uint8_t *b) {
    if (h->pos >= h->end) return 1;
    *b = h->p[h->pos++];
    return 0;
}
/* get_multibyte, xz.rs:448-464 (reader = header window) */
static int h_multibyte(hdr_reader *h, uint64_t *out, lzo_error *e) {
    uint64_t result = 0;
    int i;
    for (i = 0; i < 9; i++) {
        uint8_t byte = 0;
        if (h_u8(h, &byte)) return IOFAIL(e);
        result ^= ((uint64_t)(byte & 0x7F)) << (i * 7);
        if ((byte & 0x80) == 0) {
            *out = result;
            return 0;
        }
    }
    return fail(e, LZO_ERR_XZ, "Invalid multi-byte encoding");
}
/* get_multibyte over a digesting view (index, xz.rs:107-137): digest covers each byte read */
static int v_multibyte_digest(view *v, uint32_t *crc, uint64_t *out, lzo_error *e) {
    uint64_t result = 0;
    int i;
    for (i = 0; i < 9; i++) {
        uint8_t byte = 0;
        if (v_u8(v, &byte)) return IOFAIL(e);
        *crc = crc32_update(*crc, &byte, 1);
        result ^= ((uint64_t)(byte & 0x7F)) << (i * 7);
        if ((byte & 0x80) == 0) {
            *out = result;
            return 0;
        }
    }
    return fail(e, LZO_ERR_XZ, "Invalid multi-byte encoding");
}

typedef struct {
    uint64_t unpadded_size, unpacked_size;
} xz_record;

typedef struct {
    size_t nfilters;
    size_t props_len[4];
    int has_packed, has_unpacked;
    uint64_t packed_size, unpacked_size;
} block_header;

/* read_block_header, xz.rs:356-446 */
static int read_block_header(hdr_reader *h, uint64_t header_size, block_header *bh, lzo_error *e) {
    uint8_t flags;
    size_t num_filters, i;
    memset(bh, 0, sizeof *bh);
    if (h_u8(h, &flags)) return IOFAIL(e);
    num_filters = (size_t)(flags & 0x03) + 1;
    if (flags & 0x3C)
        return fail(e, LZO_ERR_XZ, "Invalid block flags %u, reserved bits (mask 0x3C) must be zero", (unsigned)flags);
    if (flags & 0x40) {
        bh->has_packed = 1;
        TRY(h_multibyte(h, &bh->packed_size, e));
    }
    if (flags & 0x80) {
        bh->has_unpacked = 1;
        TRY(h_multibyte(h, &bh->unpacked_size, e));
    }
    for (i = 0; i < num_filters; i++) {
        uint64_t id = 0, psize = 0;
        size_t avail;
        TRY(h_multibyte(h, &id, e));
        if (id != 0x21) return fail(e, LZO_ERR_XZ, "Unknown filter id %llu", (unsigned long long)id); /* 178-183 */
        TRY(h_multibyte(h, &psize, e));
        if (psize > header_size)
            return fail(e, LZO_ERR_XZ, "Size of filter properties exceeds block header size (%llu > %llu)",
                        (unsigned long long)psize, (unsigned long long)header_size);
        avail = h->end - h->pos;
        if (avail < psize) {
            h->pos = h->end;
            return fail(e, LZO_ERR_XZ, "Could not read filter properties of size %llu: " IO_EOF_MSG,
                        (unsigned long long)psize);
        }
        h->pos += (size_t)psize;
        bh->props_len[i] = (size_t)psize;
    }
    bh->nfilters = num_filters;
    /* flush_zero_padding, util.rs:14-34: everything left in the header must be zero */
    while (h->pos < h->end) {
        if (h->p[h->pos] != 0)
            return fail(e, LZO_ERR_XZ, "Invalid block header padding, must be null bytes");
        h->pos++;
    }
    return 0;
}

/* read_block, xz.rs:196-290.  `start` = reader position of the header_size byte. */
static int xz_read_block(reader *r, size_t start, bytes *sink, int check, xz_record **records, size_t *nrec,
                         uint8_t header_size_byte, lzo_error *e) {
    uint32_t crc = 0xFFFFFFFFu, crc_read, digest;
    uint64_t header_size = ((uint64_t)header_size_byte << 2) - 1;
    view all = view_all(r);
    hdr_reader h;
    block_header bh;
    bytes tmpbuf = {0, 0, 0};
    size_t i, count, padding_size, unpacked_size;
    int rc;

    crc = crc32_update(crc, &header_size_byte, 1);
    /* Take(header_size) pulled in one go by the BufReader: digested whether or not parsed */
    h.p = r->p;
    h.pos = r->pos;
    h.end = (header_size > (uint64_t)(r->end - r->pos)) ? r->end : r->pos + (size_t)header_size;
    crc = crc32_update(crc, r->p + r->pos, h.end - r->pos);
    r->pos = h.end;
    TRY(read_block_header(&h, header_size, &bh, e));

    if (v_u32le(&all, &crc_read)) return IOFAIL(e);
    digest = crc ^ 0xFFFFFFFFu;
    if (crc_read != digest)
        return fail(e, LZO_ERR_XZ, "Invalid header CRC32: expected 0x%08x but got 0x%08x", crc_read, digest);

    for (i = 0; i < bh.nfilters; i++) { /* xz.rs:226-250 */
        if (bh.props_len[i] != 1) { /* decode_filter, xz.rs:343-348 */
            free(tmpbuf.data);
            return fail(e, LZO_ERR_XZ, "Invalid properties for filter Lzma2");
        }
        if (i == 0) {
            size_t before = r->pos, packed;
            view in = view_all(r);
            rc = lzma2_decompress_view(&in, &tmpbuf, e);
            if (rc) {
                free(tmpbuf.data);
                return rc;
            }
            packed = r->pos - before;
            if (bh.has_packed && (uint64_t)packed != bh.packed_size) {
                free(tmpbuf.data);
                return fail(e, LZO_ERR_XZ, "Invalid compressed size: expected %llu but got %zu",
                            (unsigned long long)bh.packed_size, packed);
            }
        } else { /* chained filter decodes the previous filter's output, xz.rs:240-249 */
            bytes newbuf = {0, 0, 0};
            reader r2;
            view in2;
            r2.p = tmpbuf.data;
            r2.pos = 0;
            r2.end = tmpbuf.len;
            in2 = view_all(&r2);
            rc = lzma2_decompress_view(&in2, &newbuf, e);
            free(tmpbuf.data);
            tmpbuf = newbuf;
            if (rc) {
                free(tmpbuf.data);
                return rc;
            }
        }
    }

    unpacked_size = tmpbuf.len;
    if (bh.has_unpacked && (uint64_t)unpacked_size != bh.unpacked_size) {
        free(tmpbuf.data);
        return fail(e, LZO_ERR_XZ, "Invalid decompressed size: expected %llu but got %zu",
                    (unsigned long long)bh.unpacked_size, unpacked_size);
    }
    count = r->pos - start; /* CountBufRead::count, xz.rs:264 */
    padding_size = ((count ^ 0x03) + 1) & 0x03;
    for (i = 0; i < padding_size; i++) {
        uint8_t byte = 0;
        if (v_u8(&all, &byte)) {
            free(tmpbuf.data);
            return IOFAIL(e);
        }
        if (byte != 0) {
            free(tmpbuf.data);
            return fail(e, LZO_ERR_XZ, "Invalid block padding, must be null bytes");
        }
    }
    /* validate_block_check, xz.rs:295-333 */
    rc = 0;
    if (check == CHECK_CRC32) {
        uint32_t c, d;
        if (v_u32le(&all, &c)) rc = IOFAIL(e);
        else if (c != (d = lzo_crc32(tmpbuf.data, tmpbuf.len)))
            rc = fail(e, LZO_ERR_XZ, "Invalid block CRC32, expected 0x%08x but got 0x%08x", c, d);
    } else if (check == CHECK_CRC64) {
        uint64_t c, d;
        if (v_u64le(&all, &c)) rc = IOFAIL(e);
        else if (c != (d = lzo_crc64(tmpbuf.data, tmpbuf.len)))
            rc = fail(e, LZO_ERR_XZ, "Invalid block CRC64, expected 0x%016llx but got 0x%016llx",
                      (unsigned long long)c, (unsigned long long)d);
    } else if (check == CHECK_SHA256) {
        rc = fail(e, LZO_ERR_XZ, "Unsupported SHA-256 checksum (not yet implemented)");
    }
    if (rc) {
        free(tmpbuf.data);
        return rc;
    }
    bytes_append(sink, tmpbuf.data, tmpbuf.len); /* xz.rs:282 */
    free(tmpbuf.data);
    *records = (xz_record *)realloc(*records, (*nrec + 1) * sizeof(xz_record));
    if (!*records) abort();
    (*records)[*nrec].unpadded_size = (uint64_t)((r->pos - start) - padding_size);
    (*records)[*nrec].unpacked_size = (uint64_t)unpacked_size;
    (*nrec)++;
    return 0;
}

/* check_index, xz.rs:96-171.  `start` = position of the 0x00 index indicator. */
static int xz_check_index(reader *r, size_t start, const xz_record *records, size_t nrec, lzo_error *e) {
    uint32_t crc = 0xFFFFFFFFu, crc_read, digest;
    uint8_t tag = 0;
    view in = view_all(r);
    uint64_t num_records;
    size_t i, count, padding_size;

    crc = crc32_update(crc, &tag, 1);
    TRY(v_multibyte_digest(&in, &crc, &num_records, e));
    if (num_records != (uint64_t)nrec)
        return fail(e, LZO_ERR_XZ, "Expected %llu records but got %zu records", (unsigned long long)num_records, nrec);
    for (i = 0; i < nrec; i++) {
        uint64_t unpadded, unpacked;
        TRY(v_multibyte_digest(&in, &crc, &unpadded, e));
        if (unpadded != records[i].unpadded_size)
            return fail(e, LZO_ERR_XZ, "Invalid index for record %zu: unpadded size (%llu) does not match index (%llu)",
                        i, (unsigned long long)records[i].unpadded_size, (unsigned long long)unpadded);
        TRY(v_multibyte_digest(&in, &crc, &unpacked, e));
        if (unpacked != records[i].unpacked_size)
            return fail(e, LZO_ERR_XZ, "Invalid index for record %zu: unpacked size (%llu) does not match index (%llu)",
                        i, (unsigned long long)records[i].unpacked_size, (unsigned long long)unpacked);
    }
    count = r->pos - start;
    padding_size = ((count ^ 0x03) + 1) & 0x03;
    for (i = 0; i < padding_size; i++) {
        uint8_t byte = 0;
        if (v_u8(&in, &byte)) return IOFAIL(e);
        crc = crc32_update(crc, &byte, 1);
        if (byte != 0) return fail(e, LZO_ERR_XZ, "Invalid index padding, must be null bytes");
    }
    digest = crc ^ 0xFFFFFFFFu;
    if (v_u32le(&in, &crc_read)) return IOFAIL(e);
    if (crc_read != digest)
        return fail(e, LZO_ERR_XZ, "Invalid index CRC32: expected 0x%08x but got 0x%08x", crc_read, digest);
    return 0;
}

/* decode_stream, xz.rs:18-94 + StreamHeader::parse, xz/header.rs:20-51 */
static int xz_decompress_impl(reader *r, bytes *sink, lzo_error *e) {
    static const uint8_t XZ_MAGIC[6] = {0xFD, 0x37, 0x7A, 0x58, 0x5A, 0x00};
    view in = view_all(r);
    uint8_t magic[6], fb[2], footer_magic[2];
    uint16_t flags_field;
    uint32_t crc_read, digest, backward_size;
    int header_check = 0, footer_check = 0;
    xz_record *records = NULL;
    size_t nrec = 0, index_size;
    int rc;

    if (v_exact(&in, magic, 6)) return IOFAIL(e);
    if (memcmp(magic, XZ_MAGIC, 6) != 0)
        return fail(e, LZO_ERR_XZ, "Invalid XZ magic, expected [253, 55, 122, 88, 90, 0]");
    if (v_exact(&in, fb, 2)) return IOFAIL(e);
    flags_field = (uint16_t)((fb[0] << 8) | fb[1]);
    digest = lzo_crc32(fb, 2);
    if (v_u32le(&in, &crc_read)) return IOFAIL(e);
    if (crc_read != digest)
        return fail(e, LZO_ERR_XZ, "Invalid header CRC32: expected 0x%08x but got 0x%08x", crc_read, digest);
    TRY(stream_flags_parse(flags_field, &header_check, e));

    for (;;) { /* xz.rs:26-45 */
        size_t start = r->pos;
        uint8_t header_size;
        if (v_u8(&in, &header_size)) {
            free(records);
            return IOFAIL(e);
        }
        if (header_size == 0) {
            rc = xz_check_index(r, start, records, nrec, e);
            if (rc) {
                free(records);
                return rc;
            }
            index_size = r->pos - start;
            break;
        }
        rc = xz_read_block(r, start, sink, header_check, &records, &nrec, header_size, e);
        if (rc) {
            free(records);
            return rc;
        }
    }
    free(records);

    /* footer, xz.rs:47-92 */
    if (v_u32le(&in, &crc_read)) return IOFAIL(e);
    {
        uint8_t fbuf[6];
        size_t got = 0;
        uint32_t crc = 0xFFFFFFFFu, expect;
        /* backward_size: digested read of 4 bytes LE */
        if (v_exact(&in, fbuf, 4)) return IOFAIL(e);
        got = 4;
        crc = crc32_update(crc, fbuf, 4);
        backward_size = ((uint32_t)fbuf[3] << 24) | ((uint32_t)fbuf[2] << 16) | ((uint32_t)fbuf[1] << 8) | fbuf[0];
        expect = (uint32_t)(backward_size + 1u) << 2; /* u32, release-mode wrap */
        if ((uint32_t)index_size != expect)
            return fail(e, LZO_ERR_XZ, "Invalid index size: expected %u but got %zu", expect, index_size);
        if (v_exact(&in, fbuf + got, 2)) return IOFAIL(e);
        crc = crc32_update(crc, fbuf + got, 2);
        flags_field = (uint16_t)((fbuf[4] << 8) | fbuf[5]);
        TRY(stream_flags_parse(flags_field, &footer_check, e));
        if (header_check != footer_check)
            return fail(e, LZO_ERR_XZ,
                        "Flags in header (StreamFlags { check_method: %s }) does not match footer (StreamFlags { "
                        "check_method: %s })",
                        check_name(header_check), check_name(footer_check));
        digest = crc ^ 0xFFFFFFFFu;
    }
    if (crc_read != digest)
        return fail(e, LZO_ERR_XZ, "Invalid footer CRC32: expected 0x%08x but got 0x%08x", crc_read, digest);
    if (v_exact(&in, footer_magic, 2)) return IOFAIL(e);
    if (footer_magic[0] != 0x59 || footer_magic[1] != 0x5A)
        return fail(e, LZO_ERR_XZ, "Invalid footer magic, expected [89, 90]");
    if (!v_eof(&in)) return fail(e, LZO_ERR_XZ, "Unexpected data after last XZ block");
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * public entry points (lib.rs:44-60, 83-88, 100-105)
 * ---------------------------------------------------------------------------------------- */
static int finish_result(lzo_result *res, reader *r, bytes *sink, int rc) {
    res->out = sink->data;
    res->out_len = sink->len;
    res->consumed = r->pos;
    if (!rc) {
        res->err.kind = LZO_OK;
        res->err.msg[0] = 0;
    }
    return rc;
}

int lzo_lzma_decompress(const uint8_t *in, size_t in_len, const lzo_options *opt, lzo_result *res) {
    reader r = {in, 0, in_len};
    bytes sink = {0, 0, 0};
    memset(res, 0, sizeof *res);
    pthread_once(&crc_once, crc_init);
    return finish_result(res, &r, &sink, lzma_decompress_impl(&r, opt, &sink, &res->err));
}
int lzo_lzma2_decompress(const uint8_t *in, size_t in_len, lzo_result *res) {
    reader r = {in, 0, in_len};
    bytes sink = {0, 0, 0};
    view v;
    memset(res, 0, sizeof *res);
    pthread_once(&crc_once, crc_init);
    v = view_all(&r);
    return finish_result(res, &r, &sink, lzma2_decompress_view(&v, &sink, &res->err));
}
int lzo_xz_decompress(const uint8_t *in, size_t in_len, lzo_result *res) {
    reader r = {in, 0, in_len};
    bytes sink = {0, 0, 0};
    memset(res, 0, sizeof *res);
    pthread_once(&crc_once, crc_init);
    return finish_result(res, &r, &sink, xz_decompress_impl(&r, &sink, &res->err));
}
/* ------------------------------------------------------------------------------------------
 * decompress::raw::{LzmaDecoder, Lzma2Decoder} as OBJECTS (feature raw_decoder): the DecoderState lives in the
 * decoder and survives from one decompress() to the next (lzma.rs:597-648, lzma2.rs:11-82); every call builds a
 * fresh output buffer.  Used to check the product's lzb_raw_* entry points.
 * ---------------------------------------------------------------------------------------- */
struct lzo_raw {
    int fmt; /* 0 LzmaDecoder, 1 Lzma2Decoder */
    decstate st;
    uint32_t lc, lp, pb, dict_size;
    size_t memlimit;
};
lzo_raw *lzo_raw_new(int fmt, uint32_t lc, uint32_t lp, uint32_t pb, uint32_t dict_size, int has_unpacked,
                     uint64_t unpacked, int has_memlimit, uint64_t memlimit) {
    lzo_raw *r = (lzo_raw *)calloc(1, sizeof *r);
    r->fmt = fmt;
    r->lc = lc;
    r->lp = lp;
    r->pb = pb;
    r->dict_size = dict_size;
    r->memlimit = has_memlimit ? (size_t)memlimit : (size_t)-1; /* lzma.rs:610 */
    if (fmt == 0)
        ds_new(&r->st, lc, lp, pb, has_unpacked, unpacked); /* LzmaDecoder::new, lzma.rs:607-613 */
    else
        ds_new(&r->st, 0, 0, 0, 0, 0); /* Lzma2Decoder::new, lzma2.rs:23-34 */
    pthread_once(&crc_once, crc_init);
    return r;
}
/* LzmaDecoder::reset(unpacked_size: Option<Option<u64>>), lzma.rs:620-627; Lzma2Decoder::reset, lzma2.rs:41-48 */
void lzo_raw_reset(lzo_raw *r, int set_unpacked, int has_unpacked, uint64_t unpacked) {
    if (r->fmt == 0) {
        ds_reset_state(&r->st, r->lc, r->lp, r->pb);
        if (set_unpacked) {
            r->st.has_unpacked = has_unpacked;
            r->st.unpacked_size = unpacked;
        }
    } else {
        ds_reset_state(&r->st, 0, 0, 0);
    }
}
int lzo_raw_decompress(lzo_raw *r, const uint8_t *in, size_t in_len, lzo_result *res) {
    reader rd = {in, 0, in_len};
    bytes sink = {0, 0, 0};
    view v = view_all(&rd);
    int rcode;
    memset(res, 0, sizeof *res);
    if (r->fmt == 0) { /* LzmaDecoder::decompress, lzma.rs:635-648 */
        lzbuf out;
        rangedec rc;
        lzb_init_circ(&out, &sink, r->dict_size, r->memlimit);
        if (rc_new(&rc, v)) {
            rcode = fail(&res->err, LZO_ERR_LZMA, "LZMA stream too short: " IO_EOF_MSG);
        } else {
            rcode = ds_process(&r->st, &out, &rc, &res->err);
            if (!rcode) lzb_finish(&out);
        }
        lzb_drop(&out);
    } else {
        rcode = lzma2_decompress_state(&v, &sink, &r->st, &res->err);
    }
    return finish_result(res, &rd, &sink, rcode);
}
void lzo_raw_free(lzo_raw *r) {
    if (!r) return;
    ds_drop(&r->st);
    free(r);
}

void lzo_result_free(lzo_result *res) {
    free(res->out);
    res->out = NULL;
    res->out_len = 0;
}

/* ------------------------------------------------------------------------------------------
 * batch driver for the CPU baseline (bench.py only): one stream per task, work-stealing
 * over an atomic counter, `nthreads` pthreads.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int fmt;
    const uint8_t *in;
    const uint64_t *in_off;
    uint32_t n;
    uint8_t *out;
    const uint64_t *out_off;
    uint64_t *out_len;
    int32_t *kinds;
    uint32_t next;
    int failed;
} batch_ctx;

static void *batch_worker(void *arg) {
    batch_ctx *c = (batch_ctx *)arg;
    for (;;) {
        uint32_t i = __atomic_fetch_add(&c->next, 1, __ATOMIC_RELAXED);
        lzo_result res;
        int rc;
        uint64_t cap;
        if (i >= c->n) break;
        tl_capacity_hint = (size_t)(c->out_off[i + 1] - c->out_off[i]) + 64;
        {
            const uint8_t *p = c->in + c->in_off[i];
            size_t len = (size_t)(c->in_off[i + 1] - c->in_off[i]);
            if (c->fmt == 0) rc = lzo_lzma_decompress(p, len, NULL, &res);
            else if (c->fmt == 1) rc = lzo_lzma2_decompress(p, len, &res);
            else rc = lzo_xz_decompress(p, len, &res);
        }
        cap = c->out_off[i + 1] - c->out_off[i];
        c->out_len[i] = res.out_len;
        if (!rc && res.out_len > cap) rc = LZO_ERR_IO;
        if (res.out_len <= cap && res.out_len) memcpy(c->out + c->out_off[i], res.out, res.out_len);
        c->kinds[i] = rc;
        if (rc) __atomic_fetch_add(&c->failed, 1, __ATOMIC_RELAXED);
        lzo_result_free(&res);
    }
    return NULL;
}

int lzo_decompress_batch(int fmt, const uint8_t *in, const uint64_t *in_off, uint32_t n, uint8_t *out,
                         const uint64_t *out_off, uint64_t *out_len, int32_t *kinds, int nthreads) {
    batch_ctx c;
    pthread_t *th;
    int t;
    if (nthreads < 1) nthreads = 1;
    pthread_once(&crc_once, crc_init);
    /* keep per-stream buffers in the per-thread malloc arenas instead of mmap/munmap (process-wide lock) */
    mallopt(M_MMAP_THRESHOLD, 1 << 28);
    mallopt(M_TRIM_THRESHOLD, 1 << 29);
    c.fmt = fmt;
    c.in = in;
    c.in_off = in_off;
    c.n = n;
    c.out = out;
    c.out_off = out_off;
    c.out_len = out_len;
    c.kinds = kinds;
    c.next = 0;
    c.failed = 0;
    th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) abort();
    for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &c);
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    return c.failed;
}
