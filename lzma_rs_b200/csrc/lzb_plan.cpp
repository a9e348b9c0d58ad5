// lzb_plan.cpp -- host-side planning: .lzma header parse, LZMA2 framing scan, XZ container walk, status text.
// Reference behaviour followed (gendx/lzma-rs @ 1f14478): src/decode/lzma.rs:96-161 (header),
// src/decode/lzma2.rs:84-229 (chunk framing), src/decode/xz.rs:18-464 + src/xz/{mod,header,footer}.rs (container),
// src/error.rs:28-37 (Display).  The bit-level work (range decoder, LZ window, CRC of decoded bytes) runs on
// the GPU through `Executor`; nothing here decodes a stream.
#include "lzb_plan.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>

namespace lzb {

// ------------------------------------------------------------------------------------------------
// CRC-32/ISO-HDLC for the container's own few-byte fields (stream header, block header, index, footer)
// ------------------------------------------------------------------------------------------------
static uint32_t g_crc32_tab[256];
static bool g_crc32_init = false;
static uint32_t crc32_update(uint32_t reg, const uint8_t* p, size_t n) {
    if (!g_crc32_init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
            g_crc32_tab[i] = c;
        }
        g_crc32_init = true;
    }
    for (size_t i = 0; i < n; i++) reg = g_crc32_tab[(reg ^ p[i]) & 0xFF] ^ (reg >> 8);
    return reg;
}
uint32_t crc32_host(const uint8_t* p, size_t n) { return crc32_update(0xFFFFFFFFu, p, n) ^ 0xFFFFFFFFu; }

// ------------------------------------------------------------------------------------------------
// status helpers
// ------------------------------------------------------------------------------------------------
static int kind_of(int code) {
    if (code == LZB_OK) return LZB_KIND_OK;
    if (code < 0) return LZB_KIND_INTERNAL;
    if (code == LZB_E_IO_EOF) return LZB_KIND_IO;
    if (code == LZB_E_HEADER_TOO_SHORT) return LZB_KIND_HEADER_TOO_SHORT;
    if (code < 30) return LZB_KIND_LZMA;
    return LZB_KIND_XZ;
}
static lzb_status mk(int code, uint64_t a0 = 0, uint64_t a1 = 0, uint64_t a2 = 0) {
    lzb_status s;
    s.code = code;
    s.kind = kind_of(code);
    s.a0 = a0;
    s.a1 = a1;
    s.a2 = a2;
    return s;
}
void status_from_result(const LzbResult& r, lzb_status* st) { *st = mk(r.code, r.a0, r.a1, 0); }

// ------------------------------------------------------------------------------------------------
// scans
// ------------------------------------------------------------------------------------------------
Lzma2Scan scan_lzma2(const uint8_t* p, uint64_t len) {
    Lzma2Scan s;
    uint64_t q = 0;
    uint32_t lclp = 0;
    for (;;) {
        if (q >= len) break;
        uint32_t status = p[q++];
        if (status == 0) {
            s.well_formed = true;
            break;
        }
        if (status == 1 || status == 2) {
            if (len - q < 2) break;
            uint64_t nb = (((uint32_t)p[q] << 8) | p[q + 1]) + 1;
            q += 2;
            if (len - q < nb) break;
            q += nb;
            s.unpacked += nb;
            s.stored += nb;
            continue;
        }
        if (status < 0x80) break;
        if (len - q < 4) break;
        uint64_t unpacked = ((((uint32_t)(status & 0x1F)) << 16) | ((uint32_t)p[q] << 8) | p[q + 1]) + 1;
        uint64_t packed = (((uint32_t)p[q + 2] << 8) | p[q + 3]) + 1;
        q += 4;
        if (status >= 0xC0) {
            if (q >= len) break;
            uint32_t props = p[q++];
            if (props >= 225) break;
            uint32_t c = props % 9, l = (props / 9) % 5;
            if (c + l > 4) break;
            lclp = c + l;
        }
        s.max_lclp = std::max(s.max_lclp, lclp);
        s.unpacked += unpacked;
        if (len - q < packed) break;
        q += packed;
    }
    s.packed = q;
    return s;
}

static void item_defaults(LzbItem* it, uint64_t in_off, uint64_t in_len) {
    memset(it, 0, sizeof *it);
    it->in_off = in_off;
    it->in_len = in_len;
    it->unpacked = LZB_UNKNOWN_SIZE;
    it->memlimit = ~0ull;
    it->kind = LZB_ITEM_LZMA2;
}
static void preset(LzbItem* it, int code, uint64_t a0 = 0) {
    it->kind = LZB_ITEM_PRESET;
    it->preset_code = code;
    it->preset_a0 = a0;
}

// LzmaParams::read_header, lzma.rs:96-161
void plan_lzma(const uint8_t* p, uint64_t len, uint64_t base_off, const lzb_options* opt, LzbItem* it, LzbScan* sc) {
    static const lzb_options defaults = {0, 0, 0, 0, {0, 0, 0, 0}, 0, 0};
    if (!opt) opt = &defaults;
    item_defaults(it, base_off, len);
    it->kind = LZB_ITEM_LZMA;
    it->memlimit = opt->has_memlimit ? opt->memlimit : ~0ull;
    memset(sc, 0, sizeof *sc);
    const uint32_t hdr = opt->unpacked_mode == LZB_UNPACKED_USE_PROVIDED ? 5u : 13u;
    if (len < 1) return preset(it, LZB_E_HEADER_TOO_SHORT);
    if (p[0] >= 225) return preset(it, LZB_E_LZMA_PROPS, p[0]);
    if (len < hdr) return preset(it, LZB_E_HEADER_TOO_SHORT);
    uint32_t props = p[0];
    it->lc = props % 9;
    props /= 9;
    it->lp = props % 5;
    it->pb = props / 5;
    uint32_t ds = (uint32_t)p[1] | ((uint32_t)p[2] << 8) | ((uint32_t)p[3] << 16) | ((uint32_t)p[4] << 24);
    it->dict_size = ds < 0x1000u ? 0x1000u : ds;
    uint64_t hv = LZB_UNKNOWN_SIZE;
    if (hdr == 13) {
        hv = 0;
        for (int k = 7; k >= 0; k--) hv = (hv << 8) | p[5 + k];
    }
    if (opt->unpacked_mode == LZB_UNPACKED_READ_FROM_HEADER)
        it->unpacked = hv;
    else
        it->unpacked = opt->has_provided ? opt->provided : LZB_UNKNOWN_SIZE;
    it->in_off += hdr;
    it->in_len -= hdr;
    it->hdr_len = hdr;
    sc->max_lclp = it->lc + it->lp;
    if (it->unpacked != LZB_UNKNOWN_SIZE) {
        sc->unpacked = it->unpacked;
        sc->flags = 1;
    } else {
        sc->unpacked = len * 8 + 4096;
    }
    if (len > 0xFFFFE000ull) preset(it, LZB_E_UNSUPPORTED);
}

void plan_lzma2(const uint8_t* p, uint64_t len, uint64_t base_off, LzbItem* it, LzbScan* sc) {
    item_defaults(it, base_off, len);
    Lzma2Scan s = scan_lzma2(p, len);
    memset(sc, 0, sizeof *sc);
    sc->unpacked = s.unpacked;
    sc->flags = (s.well_formed ? 1u : 0u) | (s.stored ? 2u : 0u);
    sc->stored = s.stored;
    if (s.well_formed && s.unpacked > 0 && s.stored == s.unpacked) {
        it->flags |= LZB_ITEM_F_ALL_STORED;
        it->unpacked = s.unpacked;
    }
    sc->max_lclp = (uint8_t)s.max_lclp;
    if (len > 0xFFFFE000ull) preset(it, LZB_E_UNSUPPORTED);
}

// ------------------------------------------------------------------------------------------------
// XZ container
// ------------------------------------------------------------------------------------------------
enum { CHECK_NONE = 0x00, CHECK_CRC32 = 0x01, CHECK_CRC64 = 0x04, CHECK_SHA256 = 0x0A };

struct Cursor {  // a bounded byte reader: `end` plays io::Take / EOF
    const uint8_t* p;
    uint64_t pos, end;
    bool u8(uint8_t* b) {
        if (pos >= end) return false;
        *b = p[pos++];
        return true;
    }
    bool le32(uint32_t* v) {
        if (end - pos < 4) {
            pos = end;
            return false;
        }
        *v = (uint32_t)p[pos] | ((uint32_t)p[pos + 1] << 8) | ((uint32_t)p[pos + 2] << 16) | ((uint32_t)p[pos + 3] << 24);
        pos += 4;
        return true;
    }
    bool le64(uint64_t* v) {
        if (end - pos < 8) {
            pos = end;
            return false;
        }
        *v = 0;
        for (int k = 7; k >= 0; k--) *v = (*v << 8) | p[pos + k];
        pos += 8;
        return true;
    }
};

// get_multibyte, xz.rs:448-464.  crc != nullptr: every byte read is digested (index, xz.rs:107).
static int multibyte(Cursor& c, uint64_t* out, uint32_t* crc) {
    uint64_t r = 0;
    for (int i = 0; i < 9; i++) {
        uint8_t b;
        if (!c.u8(&b)) return LZB_E_IO_EOF;
        if (crc) *crc = crc32_update(*crc, &b, 1);
        r ^= ((uint64_t)(b & 0x7F)) << (i * 7);
        if ((b & 0x80) == 0) {
            *out = r;
            return LZB_OK;
        }
    }
    return LZB_E_XZ_MULTIBYTE;
}

// StreamFlags::parse + CheckMethod::try_from, xz/mod.rs:24-39, 62-76
static lzb_status parse_stream_flags(uint8_t b0, uint8_t b1, int* check) {
    if (b0 != 0) return mk(LZB_E_XZ_FLAGS_NULL, b0);
    if (b1 != CHECK_NONE && b1 != CHECK_CRC32 && b1 != CHECK_CRC64 && b1 != CHECK_SHA256)
        return mk(LZB_E_XZ_CHECK_METHOD, b1);
    *check = b1;
    return mk(LZB_OK);
}

struct BlockHeader {
    uint64_t start = 0;    // position of the header-size byte
    uint64_t payload = 0;  // first byte of the LZMA2 data
    bool has_packed = false, has_unpacked = false;
    uint64_t packed = 0, unpacked = 0;
    uint32_t nfilters = 0;
    uint32_t bad_props_from = 0;  // first filter index >= 1 whose properties are not 1 byte (0 = none)
};

// read_block (header part, xz.rs:206-224) + read_block_header (356-446) + the filter-props check of
// decode_filter (343-348).  On success `bh` is filled and status is OK.
static lzb_status parse_block_header(const uint8_t* p, uint64_t len, uint64_t pos, BlockHeader* bh) {
    const uint8_t hs_byte = p[pos];
    const uint64_t header_size = ((uint64_t)hs_byte << 2) - 1;
    Cursor c{p, pos + 1, std::min<uint64_t>(len, pos + 1 + header_size)};  // io::Take(header_size)
    // the BufReader over the Take pulls (and digests) the whole window before anything is parsed
    uint32_t crc = crc32_update(0xFFFFFFFFu, &hs_byte, 1);
    crc = crc32_update(crc, p + c.pos, c.end - c.pos);
    uint8_t flags;
    size_t props_len[4] = {0, 0, 0, 0};
    bh->start = pos;
    if (!c.u8(&flags)) return mk(LZB_E_IO_EOF);
    bh->nfilters = (flags & 0x03u) + 1;
    if (flags & 0x3C) return mk(LZB_E_XZ_BLOCK_FLAGS, flags);
    if (flags & 0x40) {
        bh->has_packed = true;
        int e = multibyte(c, &bh->packed, nullptr);
        if (e) return mk(e);
    }
    if (flags & 0x80) {
        bh->has_unpacked = true;
        int e = multibyte(c, &bh->unpacked, nullptr);
        if (e) return mk(e);
    }
    for (uint32_t i = 0; i < bh->nfilters; i++) {
        uint64_t id, psize;
        int e = multibyte(c, &id, nullptr);
        if (e) return mk(e);
        if (id != 0x21) return mk(LZB_E_XZ_FILTER_ID, id);
        e = multibyte(c, &psize, nullptr);
        if (e) return mk(e);
        if (psize > header_size) return mk(LZB_E_XZ_PROPS_SIZE, psize, header_size);
        if (c.end - c.pos < psize) return mk(LZB_E_XZ_PROPS_READ, psize);
        c.pos += psize;
        props_len[i] = (size_t)psize;
    }
    for (; c.pos < c.end; c.pos++)  // flush_zero_padding, util.rs:14-34
        if (p[c.pos] != 0) return mk(LZB_E_XZ_HEADER_PADDING);
    Cursor f{p, c.end, len};
    uint32_t crc_read;
    if (!f.le32(&crc_read)) return mk(LZB_E_IO_EOF);
    const uint32_t digest = crc ^ 0xFFFFFFFFu;
    if (crc_read != digest) return mk(LZB_E_XZ_HEADER_CRC, crc_read, digest);
    if (props_len[0] != 1) return mk(LZB_E_XZ_FILTER_PROPS);
    for (uint32_t i = 1; i < bh->nfilters; i++)  // decode_filter checks each filter's props when it runs (xz.rs:343-348);
        if (props_len[i] != 1) bh->bad_props_from = bh->bad_props_from ? bh->bad_props_from : i;  // later ones fail later
    bh->payload = f.pos;
    return mk(LZB_OK);
}

struct Record {
    uint64_t unpadded, unpacked;
};

// check_index (xz.rs:96-171) + footer (xz.rs:47-92).  pos = position of the 0x00 index indicator.
static lzb_status check_index_and_footer(const uint8_t* p, uint64_t len, uint64_t pos, const std::vector<Record>& records,
                                         int header_check) {
    Cursor c{p, pos + 1, len};
    const uint8_t tag = 0;
    uint32_t crc = crc32_update(0xFFFFFFFFu, &tag, 1);
    uint64_t num;
    int e = multibyte(c, &num, &crc);
    if (e) return mk(e);
    if (num != records.size()) return mk(LZB_E_XZ_INDEX_COUNT, num, records.size());
    for (size_t i = 0; i < records.size(); i++) {
        uint64_t v;
        if ((e = multibyte(c, &v, &crc))) return mk(e);
        if (v != records[i].unpadded) return mk(LZB_E_XZ_INDEX_UNPADDED, i, records[i].unpadded, v);
        if ((e = multibyte(c, &v, &crc))) return mk(e);
        if (v != records[i].unpacked) return mk(LZB_E_XZ_INDEX_UNPACKED, i, records[i].unpacked, v);
    }
    const uint64_t count = c.pos - pos;
    const uint64_t pad = ((count ^ 3) + 1) & 3;
    for (uint64_t i = 0; i < pad; i++) {
        uint8_t b;
        if (!c.u8(&b)) return mk(LZB_E_IO_EOF);
        crc = crc32_update(crc, &b, 1);
        if (b != 0) return mk(LZB_E_XZ_INDEX_PADDING);
    }
    uint32_t crc_read;
    const uint32_t digest = crc ^ 0xFFFFFFFFu;
    if (!c.le32(&crc_read)) return mk(LZB_E_IO_EOF);
    if (crc_read != digest) return mk(LZB_E_XZ_INDEX_CRC, crc_read, digest);
    const uint64_t index_size = c.pos - pos;

    // footer
    uint32_t fcrc;
    if (!c.le32(&fcrc)) return mk(LZB_E_IO_EOF);
    if (c.end - c.pos < 4) return mk(LZB_E_IO_EOF);
    const uint8_t* fb = p + c.pos;
    uint32_t backward;
    c.le32(&backward);
    const uint32_t expect = (uint32_t)(backward + 1u) << 2;  // u32 arithmetic as in xz.rs:52
    if ((uint32_t)index_size != expect) return mk(LZB_E_XZ_INDEX_SIZE, expect, index_size);
    if (c.end - c.pos < 2) return mk(LZB_E_IO_EOF);
    uint8_t f0 = p[c.pos], f1 = p[c.pos + 1];
    c.pos += 2;
    int footer_check = 0;
    lzb_status s = parse_stream_flags(f0, f1, &footer_check);
    if (s.code) return s;
    if (footer_check != header_check) return mk(LZB_E_XZ_FLAGS_MISMATCH, header_check, footer_check);
    const uint32_t fdigest = crc32_host(fb, 6);
    if (fcrc != fdigest) return mk(LZB_E_XZ_FOOTER_CRC, fcrc, fdigest);
    if (c.end - c.pos < 2) return mk(LZB_E_IO_EOF);
    if (p[c.pos] != 0x59 || p[c.pos + 1] != 0x5A) return mk(LZB_E_XZ_FOOTER_MAGIC);
    c.pos += 2;
    if (c.pos != c.end) return mk(LZB_E_XZ_TRAILING_DATA);
    return mk(LZB_OK);
}

// StreamHeader::parse, xz/header.rs:20-51
static lzb_status parse_stream_header(const uint8_t* p, uint64_t len, int* check, uint64_t* pos) {
    static const uint8_t MAGIC[6] = {0xFD, 0x37, 0x7A, 0x58, 0x5A, 0x00};
    if (len < 6) return mk(LZB_E_IO_EOF);
    if (memcmp(p, MAGIC, 6) != 0) return mk(LZB_E_XZ_MAGIC);
    if (len < 8) return mk(LZB_E_IO_EOF);
    const uint32_t digest = crc32_host(p + 6, 2);
    Cursor c{p, 8, len};
    uint32_t crc_read;
    if (!c.le32(&crc_read)) return mk(LZB_E_IO_EOF);
    if (crc_read != digest) return mk(LZB_E_XZ_HEADER_CRC, crc_read, digest);
    lzb_status s = parse_stream_flags(p[6], p[7], check);
    if (s.code) return s;
    *pos = c.pos;
    return mk(LZB_OK);
}

namespace {
struct BlockPlan {
    BlockHeader bh;
    uint64_t pred_packed = 0, pred_unpacked = 0;
    uint64_t out_rel = 0;  // output offset relative to the file's region
    uint64_t cap = 0;
    bool cap_is_prediction = false;
    uint32_t item = 0;
    int32_t crc_idx = -1;
    bool chain_stage = false;  // this item is one stage of a chained-filter block (not the whole block)
    bool chain_last = false;   // ... and it is the last filter (its output is the block's output)
};
enum Terminal { T_NONE, T_INDEX, T_ERROR };
struct XzFile {
    const uint8_t* p = nullptr;
    uint64_t len = 0, in_base = 0, out_base = 0, out_cap = 0;
    int check = 0;
    uint64_t pos = 0, out_pos = 0;
    std::vector<Record> records;
    bool done = false, lookahead = true;
    std::vector<BlockPlan> plan;
    Terminal terminal = T_NONE;
    lzb_status terminal_status{};
    uint64_t index_pos = 0;
    StreamOut* out = nullptr;
    // chained filters (xz.rs:240-249): the block at `pos` is decoded one filter per round
    uint32_t chain_next = 0;      // next filter to run (0 = not inside a chain)
    uint64_t chain_in_off = 0, chain_in_len = 0;  // previous filter's output (output-blob coordinates)
    uint64_t chain_packed = 0;    // bytes filter 0 consumed from the file
    std::vector<uint8_t> chain_host;  // host copy of the previous filter's output (framing scan only)
    uint64_t chain_scratch_hint = 0;  // scratch size to try after a stage overflowed its scratch
    int finish(const lzb_status& s) {
        out->st = s;
        out->out_len = out_pos;
        out->consumed = s.code == LZB_OK ? len : pos;
        done = true;
        return LZB_RC_OK;
    }
};
}  // namespace

static uint64_t check_len(int check) {
    return check == CHECK_CRC32 ? 4 : check == CHECK_CRC64 ? 8 : check == CHECK_SHA256 ? 32 : 0;
}

// One stage of a chained-filter block (xz.rs:240-249): filter j decodes the previous filter's output.  Stage 0 reads
// the file payload; every stage but the last writes to device scratch; the last writes to the file's output region.
static int plan_chain_stage(Executor& ex, XzFile& f, const BlockPlan& proto, std::vector<LzbItem>& items,
                            uint32_t* max_lclp, uint64_t* stored) {
    BlockPlan bp = proto;
    const uint32_t j = f.chain_next;
    const bool last = j + 1 == bp.bh.nfilters;
    if (bp.bh.bad_props_from && j >= bp.bh.bad_props_from) {  // decode_filter(j): props.len() != 1, xz.rs:343-348
        f.terminal = T_ERROR;
        f.terminal_status = mk(LZB_E_XZ_FILTER_PROPS);
        return LZB_RC_OK;
    }
    const uint8_t* src = j == 0 ? f.p + bp.bh.payload : f.chain_host.data();
    const uint64_t src_len = j == 0 ? f.len - bp.bh.payload : f.chain_in_len;
    Lzma2Scan sc = scan_lzma2(src, src_len);
    *max_lclp = std::max(*max_lclp, sc.max_lclp);
    *stored += sc.stored;
    LzbItem it;
    if (j == 0) {
        item_defaults(&it, f.in_base + bp.bh.payload, src_len);
    } else {
        item_defaults(&it, f.chain_in_off, src_len);
        it.flags = LZB_ITEM_F_IN_FROM_OUT;
    }
    bp.out_rel = f.out_pos;
    if (last) {
        bp.cap = f.out_cap - std::min(f.out_cap, f.out_pos);
        it.out_off = f.out_base + f.out_pos;
    } else {
        bp.cap = std::max<uint64_t>(std::max<uint64_t>(sc.unpacked, f.chain_scratch_hint), 4096) + 64;
        int rc = ex.scratch(bp.cap, &it.out_off);
        if (rc != LZB_RC_OK) return rc;
        it.flags |= LZB_ITEM_F_OUT_SCRATCH;
    }
    it.out_cap = bp.cap;
    if (it.in_len > 0xFFFFE000ull) preset(&it, LZB_E_UNSUPPORTED);
    bp.chain_stage = true;
    bp.chain_last = last;
    bp.item = (uint32_t)items.size();
    items.push_back(it);
    f.plan.push_back(bp);
    return LZB_RC_OK;
}

static int plan_file(Executor& ex, XzFile& f, std::vector<LzbItem>& items, uint32_t* max_lclp, uint64_t* stored) {
    uint64_t pos = f.pos, out_rel = f.out_pos;
    f.plan.clear();
    f.terminal = T_NONE;
    for (;;) {
        if (pos >= f.len) {  // count_input.read_u8()? at xz.rs:28
            f.terminal = T_ERROR;
            f.terminal_status = mk(LZB_E_IO_EOF);
            break;
        }
        if (f.p[pos] == 0) {
            f.terminal = T_INDEX;
            f.index_pos = pos;
            break;
        }
        BlockPlan bp;
        lzb_status s = parse_block_header(f.p, f.len, pos, &bp.bh);
        if (s.code) {
            f.terminal = T_ERROR;
            f.terminal_status = s;
            break;
        }
        if (bp.bh.nfilters > 1) {  // chained filters: the block runs alone, one filter per round
            if (!f.plan.empty()) break;
            return plan_chain_stage(ex, f, bp, items, max_lclp, stored);
        }
        Lzma2Scan sc = scan_lzma2(f.p + bp.bh.payload, f.len - bp.bh.payload);
        *max_lclp = std::max(*max_lclp, sc.max_lclp);
        *stored += sc.stored;
        const bool all_stored = sc.well_formed && sc.unpacked > 0 && sc.stored == sc.unpacked;
        bp.pred_packed = sc.packed;
        bp.pred_unpacked = sc.unpacked;
        bp.out_rel = out_rel;
        const uint64_t remaining = f.out_cap - std::min(f.out_cap, out_rel);
        const bool predict = f.lookahead && sc.well_formed && f.check != CHECK_SHA256;
        bp.cap = remaining;
        if (predict && sc.unpacked < remaining) {
            bp.cap = sc.unpacked;
            bp.cap_is_prediction = true;
        }
        LzbItem it;
        item_defaults(&it, f.in_base + bp.bh.payload, f.len - bp.bh.payload);
        it.out_off = f.out_base + out_rel;
        it.out_cap = bp.cap;
        if (all_stored) {  // what the reference's own xz_compress writes: eligible for the stored-chunk copy kernel
            it.flags |= LZB_ITEM_F_ALL_STORED;
            it.unpacked = sc.unpacked;
        }
        if (it.in_len > 0xFFFFE000ull) preset(&it, LZB_E_UNSUPPORTED);
        bp.item = (uint32_t)items.size();
        items.push_back(it);
        f.plan.push_back(bp);
        if (!predict) break;  // cannot know where the next block starts before this one is decoded
        const uint64_t end = bp.bh.payload + sc.packed;
        const uint64_t count = end - pos;
        pos = end + (((count ^ 3) + 1) & 3) + check_len(f.check);
        out_rel += sc.unpacked;
    }
    return LZB_RC_OK;
}

// Validates the planned blocks of one file in order (xz.rs:232-287); returns true when the file is finished.
static int validate_file(Executor& ex, XzFile& f, const std::vector<LzbItem>& items, const std::vector<LzbResult>& results,
                         const std::vector<uint32_t>& crc32, const std::vector<uint64_t>& crc64) {
    for (size_t k = 0; k < f.plan.size(); k++) {
        const BlockPlan& bp = f.plan[k];
        const LzbResult& r = results[bp.item];
        const uint64_t remaining = f.out_cap - std::min(f.out_cap, bp.out_rel);
        if (bp.chain_stage && !bp.chain_last) {  // an inner filter of a chain: its output feeds the next round
            if (r.code == LZB_E_CAPACITY) {  // scratch too small (framing scan mispredicted): retry this stage bigger
                f.chain_scratch_hint = std::max<uint64_t>(r.a0 * 2, f.chain_scratch_hint * 2);
                if (f.chain_scratch_hint > 0xF0000000ull) return f.finish(mk(LZB_E_UNSUPPORTED)), LZB_RC_OK;
                return LZB_RC_OK;
            }
            if (r.code != LZB_OK) return f.finish(mk(r.code, r.a0, r.a1)), LZB_RC_OK;
            if (f.chain_next == 0) {
                f.chain_packed = r.consumed;
                if (bp.bh.has_packed && r.consumed != bp.bh.packed)  // checked right after filter 0, xz.rs:232-238
                    return f.finish(mk(LZB_E_XZ_PACKED_SIZE, bp.bh.packed, r.consumed)), LZB_RC_OK;
            }
            f.chain_in_off = items[bp.item].out_off;
            f.chain_in_len = r.out_len;
            f.chain_host.resize(r.out_len + 16);
            if (r.out_len) {
                int rc = ex.read_out(f.chain_in_off, r.out_len, f.chain_host.data());
                if (rc != LZB_RC_OK) return rc;
            }
            f.chain_next += 1;
            f.chain_scratch_hint = 0;
            return LZB_RC_OK;
        }
        if (r.code < 0) {  // LZB_KIND_INTERNAL
            if (r.code == LZB_E_CAPACITY && bp.cap_is_prediction && bp.cap < remaining) {
                f.lookahead = false;  // the framing scan mispredicted: redo from this block without look-ahead
                return LZB_RC_OK;
            }
            return f.finish(mk(r.code, r.code == LZB_E_CAPACITY ? bp.out_rel + r.a0 : r.a0, r.a1));
        }
        if (r.code != LZB_OK) return f.finish(mk(r.code, r.a0, r.a1));  // `?` on Lzma2Decoder::decompress, xz.rs:350
        // (a chain's last filter: the bytes consumed from the FILE are those of filter 0)
        const uint64_t packed = (bp.chain_stage && f.chain_next > 0) ? f.chain_packed : r.consumed, unpacked = r.out_len;
        if (bp.chain_stage) f.chain_next = 0;
        if (bp.bh.has_packed && packed != bp.bh.packed) return f.finish(mk(LZB_E_XZ_PACKED_SIZE, bp.bh.packed, packed));
        if (bp.bh.has_unpacked && unpacked != bp.bh.unpacked)
            return f.finish(mk(LZB_E_XZ_UNPACKED_SIZE, bp.bh.unpacked, unpacked));
        Cursor c{f.p, bp.bh.payload + packed, f.len};
        const uint64_t count = c.pos - bp.bh.start;
        const uint64_t pad = ((count ^ 3) + 1) & 3;
        for (uint64_t i = 0; i < pad; i++) {
            uint8_t b;
            if (!c.u8(&b)) return f.finish(mk(LZB_E_IO_EOF));
            if (b != 0) return f.finish(mk(LZB_E_XZ_BLOCK_PADDING));
        }
        if (f.check == CHECK_CRC32) {  // validate_block_check, xz.rs:295-333
            uint32_t v;
            if (!c.le32(&v)) return f.finish(mk(LZB_E_IO_EOF));
            if (v != crc32[bp.crc_idx]) return f.finish(mk(LZB_E_XZ_BLOCK_CRC32, v, crc32[bp.crc_idx]));
        } else if (f.check == CHECK_CRC64) {
            uint64_t v;
            if (!c.le64(&v)) return f.finish(mk(LZB_E_IO_EOF));
            if (v != crc64[bp.crc_idx]) return f.finish(mk(LZB_E_XZ_BLOCK_CRC64, v, crc64[bp.crc_idx]));
        } else if (f.check == CHECK_SHA256) {
            return f.finish(mk(LZB_E_XZ_SHA256));
        }
        // the block is valid: its bytes count as written (xz.rs:282-286)
        f.out_pos = bp.out_rel + unpacked;
        f.records.push_back(Record{(c.pos - bp.bh.start) - pad, unpacked});
        f.pos = c.pos;
        if ((packed != bp.pred_packed || unpacked != bp.pred_unpacked) && (k + 1 < f.plan.size() || f.terminal != T_NONE)) {
            // everything planned behind this block -- later blocks, but also the index or a header error found at the
            // predicted position -- was read at the wrong offset: plan again from the real end of this block
            f.lookahead = false;
            return LZB_RC_OK;
        }
    }
    if (f.terminal == T_ERROR) return f.finish(f.terminal_status);
    if (f.terminal == T_INDEX) return f.finish(check_index_and_footer(f.p, f.len, f.index_pos, f.records, f.check));
    // T_NONE: planning stopped for lack of look-ahead; continue next round from f.pos
    return LZB_RC_OK;
}

int decode_xz_batch(Executor& ex, const uint8_t* in, const uint64_t* in_off, uint32_t n, const uint64_t* out_off,
                    StreamOut* outs) {
    std::vector<XzFile> files(n);
    for (uint32_t i = 0; i < n; i++) {
        XzFile& f = files[i];
        f.p = in + in_off[i];
        f.len = in_off[i + 1] - in_off[i];
        f.in_base = in_off[i];
        f.out_base = out_off[i];
        f.out_cap = out_off[i + 1] - out_off[i];
        f.out = &outs[i];
        lzb_status s = parse_stream_header(f.p, f.len, &f.check, &f.pos);
        if (s.code) {
            f.pos = 0;
            f.finish(s);
        }
    }
    std::vector<LzbItem> items;
    std::vector<LzbResult> results;
    std::vector<CrcRange> ranges;
    std::vector<uint32_t> c32;
    std::vector<uint64_t> c64;
    for (;;) {
        items.clear();
        uint32_t max_lclp = 0;
        uint64_t stored = 0;
        bool any = false;
        for (auto& f : files) {
            if (f.done) continue;
            any = true;
            int prc = plan_file(ex, f, items, &max_lclp, &stored);
            if (prc != LZB_RC_OK) return prc;
        }
        if (!any) break;
        results.assign(items.size(), LzbResult{});
        if (!items.empty()) {
            int rc = ex.decode(items.data(), (uint32_t)items.size(), max_lclp, stored, results.data());
            if (rc != LZB_RC_OK) return rc;
        }
        ranges.clear();
        for (auto& f : files) {
            if (f.done || (f.check != CHECK_CRC32 && f.check != CHECK_CRC64)) continue;
            for (auto& bp : f.plan) {
                const LzbResult& r = results[bp.item];
                if (r.code != LZB_OK) break;
                if (bp.chain_stage && !bp.chain_last) break;  // intermediate result of a filter chain: no check
                bp.crc_idx = (int32_t)ranges.size();
                ranges.push_back(CrcRange{f.out_base + bp.out_rel, r.out_len});
            }
        }
        c32.assign(ranges.size(), 0);
        c64.assign(ranges.size(), 0);
        if (!ranges.empty()) {
            int rc = ex.crc(ranges.data(), (uint32_t)ranges.size(), c32.data(), c64.data());
            if (rc != LZB_RC_OK) return rc;
        }
        for (auto& f : files) {
            if (f.done) continue;
            int vrc = validate_file(ex, f, items, results, c32, c64);
            if (vrc != LZB_RC_OK) return vrc;
        }
    }
    return LZB_RC_OK;
}

// Options::allow_incomplete (options.rs:15-19, honoured by stream.rs:136-147): input that ends inside a symbol ends the
// .lzma stream without an error, and everything decoded from complete symbols counts as output.  K1 reports the end
// of input at the symbol boundary, before any side effect of the unfinished symbol, which is exactly what the
// reference's dry run (lzma.rs:408-419, 470-483) leaves in the window.
// The stream API never runs the final size check either when the flag is set (it belongs to ProcessingMode::Finish,
// lzma.rs:513-521, which finish() skips): a stream that ends short of, or overshoots, its declared size also counts.
bool lenient_eof(int fmt, const lzb_options* opt, const LzbResult* r) {
    return fmt == LZB_FMT_LZMA && opt && opt->allow_incomplete &&
           (r->code == LZB_E_IO_EOF || r->code == LZB_E_UNPACKED_MISMATCH);
}

// lzma_decompress[_with_options] / lzma2_decompress over a batch (lib.rs:44-60, 83-88): one work item per stream.
int decode_batch(Executor& ex, int fmt, const lzb_options* opt, const uint8_t* in, const uint64_t* in_off, uint32_t n,
                 const uint64_t* out_off, StreamOut* outs) {
    if (fmt == LZB_FMT_XZ) return decode_xz_batch(ex, in, in_off, n, out_off, outs);
    std::vector<LzbItem> items(n);
    std::vector<LzbResult> results(n);
    uint32_t max_lclp = 0;
        uint64_t stored = 0;
    for (uint32_t i = 0; i < n; i++) {
        LzbScan sc;
        const uint8_t* p = in + in_off[i];
        const uint64_t len = in_off[i + 1] - in_off[i];
        if (fmt == LZB_FMT_LZMA)
            plan_lzma(p, len, in_off[i], opt, &items[i], &sc);
        else
            plan_lzma2(p, len, in_off[i], &items[i], &sc);
        items[i].out_off = out_off[i];
        items[i].out_cap = out_off[i + 1] - out_off[i];
        if (items[i].kind != LZB_ITEM_PRESET) max_lclp = std::max<uint32_t>(max_lclp, sc.max_lclp);
        stored += sc.stored;
    }
    if (n) {
        int rc = ex.decode(items.data(), n, max_lclp, stored, results.data());
        if (rc != LZB_RC_OK) return rc;
    }
    for (uint32_t i = 0; i < n; i++) {
        if (lenient_eof(fmt, opt, &results[i])) results[i].code = LZB_OK, results[i].sink_len = results[i].out_len;
        status_from_result(results[i], &outs[i].st);
        outs[i].out_len = results[i].sink_len;
        outs[i].consumed = items[i].hdr_len + results[i].consumed;
    }
    return LZB_RC_OK;
}

// Output capacity a stream needs (lzb_scan): exact for well-formed LZMA2 / XZ / known-size .lzma (+ slack for the
// <= 272 bytes a final match may overshoot before the size check fires, lzma.rs:513-521).
uint64_t scan_capacity(int fmt, const lzb_options* opt, const uint8_t* p, uint64_t len) {
    // sizes claimed by (possibly corrupt) headers are trusted up to a generous expansion bound; a stream that really
    // expands more reports LZB_E_CAPACITY with the bytes it needs and is retried (lzb_decompress_alloc)
    const uint64_t bound = len * 16384 + (1u << 20);
    if (fmt == LZB_FMT_XZ) return std::min(scan_xz_capacity(p, len), bound);
    LzbItem it;
    LzbScan sc;
    if (fmt == LZB_FMT_LZMA) {
        plan_lzma(p, len, 0, opt, &it, &sc);
        if (it.kind == LZB_ITEM_PRESET) return 0;
        // a declared size is trusted up to a generous expansion bound; beyond it (garbage headers) start from a
        // heuristic and let LZB_E_CAPACITY drive the retry (lzb_decompress_alloc)
        if ((sc.flags & 1) && sc.unpacked <= bound) return sc.unpacked + 288;
        return len * 8 + 65536;
    }
    plan_lzma2(p, len, 0, &it, &sc);
    return std::min<uint64_t>(sc.unpacked, bound);
}

uint64_t scan_xz_capacity(const uint8_t* p, uint64_t len) {
    int check = 0;
    uint64_t pos = 0, total = 0;
    if (parse_stream_header(p, len, &check, &pos).code) return 0;
    for (;;) {
        if (pos >= len || p[pos] == 0) break;
        BlockHeader bh;
        if (parse_block_header(p, len, pos, &bh).code) break;
        Lzma2Scan sc = scan_lzma2(p + bh.payload, len - bh.payload);
        const uint64_t cap_bound = len * 16384 + (1u << 20);  // ignore absurd sizes in corrupt block headers
        total += std::min(std::max<uint64_t>(sc.unpacked, bh.has_unpacked ? bh.unpacked : 0), cap_bound);
        if (!sc.well_formed) break;
        const uint64_t end = bh.payload + sc.packed;
        const uint64_t count = end - pos;
        pos = end + (((count ^ 3) + 1) & 3) + check_len(check);
    }
    return total;
}

}  // namespace lzb

// ------------------------------------------------------------------------------------------------
// Display strings, identical to the reference's (src/error.rs:28-37 + each format!() site)
// ------------------------------------------------------------------------------------------------
static const char* check_name(uint64_t m) {  // #[derive(Debug)] CheckMethod, xz/mod.rs:53-60
    return m == 0x00 ? "None" : m == 0x01 ? "Crc32" : m == 0x04 ? "Crc64" : "Sha256";
}

extern "C" size_t lzb_format_error(const lzb_status* st, char* buf, size_t buf_len) {
    char m[400];
    const unsigned long long a0 = st->a0, a1 = st->a1, a2 = st->a2;
    const char* eof = "failed to fill whole buffer";  // std::io::Read::read_exact at EOF
    const char* pfx = "";
    switch (st->kind) {
    case LZB_KIND_IO: pfx = "io error: "; break;
    case LZB_KIND_HEADER_TOO_SHORT: pfx = "header too short: "; break;
    case LZB_KIND_LZMA: pfx = "lzma error: "; break;
    case LZB_KIND_XZ: pfx = "xz error: "; break;
    case LZB_KIND_INTERNAL: pfx = "lzma_b200: "; break;
    default: break;
    }
    m[0] = 0;
    switch (st->code) {
    case LZB_OK: break;
    case LZB_E_IO_EOF:
    case LZB_E_HEADER_TOO_SHORT: snprintf(m, sizeof m, "%s", eof); break;
    case LZB_E_LZMA_PROPS: snprintf(m, sizeof m, "LZMA header invalid properties: %llu must be < 225", a0); break;
    case LZB_E_LZMA_STREAM_TOO_SHORT: snprintf(m, sizeof m, "LZMA stream too short: %s", eof); break;
    case LZB_E_EOS_MORE_BYTES: snprintf(m, sizeof m, "Found end-of-stream marker but more bytes are available"); break;
    case LZB_E_UNPACKED_MISMATCH: snprintf(m, sizeof m, "Expected unpacked size of %llu but decompressed to %llu", a0, a1); break;
    case LZB_E_MATCH_DIST_DICT: snprintf(m, sizeof m, "Match distance %llu is beyond dictionary size %llu", a0, a1); break;
    case LZB_E_MATCH_DIST_OUT: snprintf(m, sizeof m, "Match distance %llu is beyond output size %llu", a0, a1); break;
    case LZB_E_LZ_DIST_DICT: snprintf(m, sizeof m, "LZ distance %llu is beyond dictionary size %llu", a0, a1); break;
    case LZB_E_LZ_DIST_OUT: snprintf(m, sizeof m, "LZ distance %llu is beyond output size %llu", a0, a1); break;
    case LZB_E_MEMLIMIT: snprintf(m, sizeof m, "exceeded memory limit of %llu", a0); break;
    case LZB_E_L2_STATUS_EOF: snprintf(m, sizeof m, "LZMA2 expected new status: %s", eof); break;
    case LZB_E_L2_INVALID_STATUS: snprintf(m, sizeof m, "LZMA2 invalid status %llu, must be 0, 1, 2 or >= 128", a0); break;
    case LZB_E_L2_UNPACKED_EOF: snprintf(m, sizeof m, "LZMA2 expected unpacked size: %s", eof); break;
    case LZB_E_L2_PACKED_EOF: snprintf(m, sizeof m, "LZMA2 expected packed size: %s", eof); break;
    case LZB_E_L2_PROPS_EOF: snprintf(m, sizeof m, "LZMA2 expected new properties: %s", eof); break;
    case LZB_E_L2_PROPS_RANGE: snprintf(m, sizeof m, "LZMA2 invalid properties: %llu must be < 225", a0); break;
    case LZB_E_L2_PROPS_LCLP: snprintf(m, sizeof m, "LZMA2 invalid properties: lc + lp (%llu + %llu) must be <= 4", a0, a1); break;
    case LZB_E_L2_STORED_EOF: snprintf(m, sizeof m, "LZMA2 expected %llu uncompressed bytes: %s", a0, eof); break;
    case LZB_E_L2_INPUT_TOO_SHORT: snprintf(m, sizeof m, "LZMA input too short: %s", eof); break;
    case LZB_E_XZ_MAGIC: snprintf(m, sizeof m, "Invalid XZ magic, expected [253, 55, 122, 88, 90, 0]"); break;
    case LZB_E_XZ_HEADER_CRC: snprintf(m, sizeof m, "Invalid header CRC32: expected 0x%08llx but got 0x%08llx", a0, a1); break;
    case LZB_E_XZ_FLAGS_NULL: snprintf(m, sizeof m, "Invalid null byte in Stream Flags: %llx", a0); break;
    case LZB_E_XZ_CHECK_METHOD: snprintf(m, sizeof m, "Invalid check method %llx, expected one of [0x00, 0x01, 0x04, 0x0A]", a0); break;
    case LZB_E_XZ_BLOCK_FLAGS: snprintf(m, sizeof m, "Invalid block flags %llu, reserved bits (mask 0x3C) must be zero", a0); break;
    case LZB_E_XZ_FILTER_ID: snprintf(m, sizeof m, "Unknown filter id %llu", a0); break;
    case LZB_E_XZ_PROPS_SIZE: snprintf(m, sizeof m, "Size of filter properties exceeds block header size (%llu > %llu)", a0, a1); break;
    case LZB_E_XZ_PROPS_READ: snprintf(m, sizeof m, "Could not read filter properties of size %llu: %s", a0, eof); break;
    case LZB_E_XZ_HEADER_PADDING: snprintf(m, sizeof m, "Invalid block header padding, must be null bytes"); break;
    case LZB_E_XZ_FILTER_PROPS: snprintf(m, sizeof m, "Invalid properties for filter Lzma2"); break;
    case LZB_E_XZ_PACKED_SIZE: snprintf(m, sizeof m, "Invalid compressed size: expected %llu but got %llu", a0, a1); break;
    case LZB_E_XZ_UNPACKED_SIZE: snprintf(m, sizeof m, "Invalid decompressed size: expected %llu but got %llu", a0, a1); break;
    case LZB_E_XZ_BLOCK_PADDING: snprintf(m, sizeof m, "Invalid block padding, must be null bytes"); break;
    case LZB_E_XZ_BLOCK_CRC32: snprintf(m, sizeof m, "Invalid block CRC32, expected 0x%08llx but got 0x%08llx", a0, a1); break;
    case LZB_E_XZ_BLOCK_CRC64: snprintf(m, sizeof m, "Invalid block CRC64, expected 0x%016llx but got 0x%016llx", a0, a1); break;
    case LZB_E_XZ_SHA256: snprintf(m, sizeof m, "Unsupported SHA-256 checksum (not yet implemented)"); break;
    case LZB_E_XZ_INDEX_COUNT: snprintf(m, sizeof m, "Expected %llu records but got %llu records", a0, a1); break;
    case LZB_E_XZ_INDEX_UNPADDED: snprintf(m, sizeof m, "Invalid index for record %llu: unpadded size (%llu) does not match index (%llu)", a0, a1, a2); break;
    case LZB_E_XZ_INDEX_UNPACKED: snprintf(m, sizeof m, "Invalid index for record %llu: unpacked size (%llu) does not match index (%llu)", a0, a1, a2); break;
    case LZB_E_XZ_INDEX_PADDING: snprintf(m, sizeof m, "Invalid index padding, must be null bytes"); break;
    case LZB_E_XZ_INDEX_CRC: snprintf(m, sizeof m, "Invalid index CRC32: expected 0x%08llx but got 0x%08llx", a0, a1); break;
    case LZB_E_XZ_MULTIBYTE: snprintf(m, sizeof m, "Invalid multi-byte encoding"); break;
    case LZB_E_XZ_INDEX_SIZE: snprintf(m, sizeof m, "Invalid index size: expected %llu but got %llu", a0, a1); break;
    case LZB_E_XZ_FLAGS_MISMATCH:
        snprintf(m, sizeof m, "Flags in header (StreamFlags { check_method: %s }) does not match footer (StreamFlags { check_method: %s })",
                 check_name(a0), check_name(a1));
        break;
    case LZB_E_XZ_FOOTER_CRC: snprintf(m, sizeof m, "Invalid footer CRC32: expected 0x%08llx but got 0x%08llx", a0, a1); break;
    case LZB_E_XZ_FOOTER_MAGIC: snprintf(m, sizeof m, "Invalid footer magic, expected [89, 90]"); break;
    case LZB_E_XZ_TRAILING_DATA: snprintf(m, sizeof m, "Unexpected data after last XZ block"); break;
    case LZB_E_CAPACITY: snprintf(m, sizeof m, "output capacity too small, need at least %llu bytes", a0); break;
    case LZB_E_UNSUPPORTED: snprintf(m, sizeof m, "stream outside the GPU path's limits"); break;
    case LZB_E_INPUT_TIMEOUT: snprintf(m, sizeof m, "input upload did not reach the device"); break;
    default: snprintf(m, sizeof m, "unknown status %d", st->code); break;
    }
    int n = snprintf(buf, buf_len, "%s%s", pfx, m);
    return n < 0 ? 0 : (size_t)n;
}
