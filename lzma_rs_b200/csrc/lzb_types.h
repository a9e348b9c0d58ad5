// lzb_types.h -- work-item / result records shared by the host side and the CUDA kernels.
#pragma once
#include <stdint.h>

#include "lzma_b200.h"

// One independent compressed stream (a .lzma payload, a raw LZMA2 stream, or one .xz block).
// Offsets are relative to the input / output blobs handed to the kernel.
struct LzbItem {
    uint64_t in_off;    // first byte the decoder reads (LZMA: the range-coder bytes after the header)
    uint64_t in_len;    // bytes readable from in_off (to the end of the caller's input for this stream)
    uint64_t out_off;   // where this stream's output starts in the output blob
    uint64_t out_cap;   // capacity in bytes
    uint64_t unpacked;  // LZMA: expected unpacked size, or LZB_UNKNOWN_SIZE (end-marker mode)
    uint64_t memlimit;  // LZMA: Options::memlimit (UINT64_MAX = unlimited)
    uint32_t dict_size; // LZMA: max(header dict size, 0x1000)   (lzma.rs:122-126)
    uint8_t kind;       // LZB_ITEM_*
    uint8_t lc, lp, pb; // LZMA properties (LZMA2 reads its own from chunk headers)
    uint32_t hdr_len;   // bytes of container header already consumed before in_off (LZMA: 13 or 5)
    int32_t preset_code;// LZB_ITEM_PRESET: status decided before the decode (header errors)
    uint32_t flags;     // LZB_ITEM_F_*
    uint64_t preset_a0;
    uint64_t host_out;  // device-visible address of this stream's region in the caller's PINNED host output (0 = none):
                        // the mirror variant of K1 streams finished 4 KiB pages there while it decodes
};

enum { LZB_ITEM_LZMA = 0, LZB_ITEM_LZMA2 = 1, LZB_ITEM_PRESET = 2 };
// in_off addresses the OUTPUT blob: the stream is the output of an earlier launch (chained .xz filters, xz.rs:240-249)
// out_off addresses device scratch outside the caller's output region (never mirrored to the host)
// the framing scan found a well-formed LZMA2 stream made of stored chunks only (what the reference's own encoders write,
// src/encode/lzma2.rs:4-26); `unpacked` then holds its size.  Such streams need no range decoder: they are routed to
// the stored-chunk copy kernel (lzb_stored_decode_kernel) when their output fits the capacity.
enum { LZB_ITEM_F_IN_FROM_OUT = 1, LZB_ITEM_F_OUT_SCRATCH = 2, LZB_ITEM_F_ALL_STORED = 4, LZB_ITEM_F_CARRY = 8 };
#define LZB_UNKNOWN_SIZE 0xFFFFFFFFFFFFFFFFull
// K1's order array: a warp whose pre-assigned first entry is this value takes no part in the launch (lzb_sched.h)
#define LZB_ORDER_PARK 0xFFFFFFFFu

// decompress::raw decoders keep their DecoderState between two decompress() calls (src/decode/lzma.rs:597-633,
// src/decode/lzma2.rs:11-48): probabilities, state, rep[4] and -- for LZMA2 -- the properties of the last props reset; the
// output window does NOT persist (every call builds a fresh LzCircularBuffer / LzAccumBuffer).  A work item with
// LZB_ITEM_F_CARRY points (host_out) at this record in device memory; the carry kernel starts from it and writes it back.
struct LzbCarry {
    uint32_t fresh;     // 1: DecoderState::new / reset_state -- every probability 0x400, state 0, rep 0
    uint32_t state;
    uint32_t rep[4];
    uint32_t lc, lp, pb;
    uint32_t lclp_cap;  // the literal area below holds 0x300 << lclp_cap entries
    uint32_t pad[6];
    // uint16_t small[T_LIT];             the small tables in K1's layout (T_IS_MATCH .. T_REP_LEN)
    // uint16_t lit[0x300 << lclp_cap];   the literal table in the reference's layout, used in place by the kernel
};
static inline uint64_t lzb_carry_bytes(uint32_t lclp_cap);

// Per-stream result written by the decode kernel.
struct LzbResult {
    int32_t code;      // LZB_OK / LZB_E_*
    uint32_t chunks;   // LZMA2 chunks walked (diagnostic)
    uint64_t a0, a1;   // message arguments of the failing site
    uint64_t out_len;  // bytes produced into the output buffer
    uint64_t sink_len; // bytes the reference would have handed to its io::Write sink
    uint64_t consumed; // input bytes consumed from in_off
};

// Per-stream scan summary (K2).
struct LzbScan {
    uint64_t unpacked; // exact for well-formed LZMA2 / known-size LZMA; heuristic otherwise
    uint32_t flags;    // bit0: walk ended at a 0x00 control byte (well-formed framing); bit1: has a stored chunk
    uint8_t max_lclp;  // largest lc+lp any chunk (or the .lzma header) asks for
    uint8_t pad[3];
    uint64_t stored;   // bytes of `unpacked` in stored chunks
};

// Probability-table layout in shared memory (u16 indices); see DESIGN.md "K1".
enum {
    T_IS_MATCH = 0,      // [12][16]  (state<<4)+pos_state        lzma.rs:175,289
    T_IS_REP = 192,      // [12]                                  lzma.rs:176
    T_IS_REP_G0 = 204,   // [12]
    T_IS_REP_G1 = 216,   // [12]
    T_IS_REP_G2 = 228,   // [12]
    T_IS_REP0LONG = 240, // [12][16]                              lzma.rs:180,317
    T_POS_SLOT = 432,    // [4][64]                               lzma.rs:172
    T_POS_DEC = 688,     // 10 reverse trees (slots 4..13), one 4-byte aligned block each: 124 entries
                         // (the reference packs them into [115] with overlapping offsets, lzma.rs:174,579-585)
    T_ALIGN = 812,       // [16]                                  lzma.rs:173
    T_LEN = 828,         // choice, choice2, low[16][8], mid[16][8], high[256]   rangecoder.rs:202-209
    T_REP_LEN = 1342,    // same
    T_LIT = 1856,        // PLAIN part of the literal table: [1<<(lc+lp)][0x100] (columns 0..0xFF of lzma.rs:194);
                         // the matched-literal columns 0x100..0x2FF live in a per-warp global workspace
    T_LEN_SIZE = 514,
    T_LEN_LOW = 2,
    T_LEN_MID = 2 + 128,
    T_LEN_HIGH = 2 + 256
};

// Multipliers / addends the bit loop reads from the kernel's constant bank.  Passing them as launch parameters keeps
// nvcc/ptxas from strength-reducing `x * 2 + y` into ALU-pipe shifts and selects: as IMADs with a constant-bank
// operand they run on the (otherwise idle) FMA pipe, which halves the ALU-pipe pressure that bounds K1.
// Warps (= resident streams) per CTA, one CTA per SM: bounded by shared memory (28 x 7 808 B of tables) and by
// registers (65 536 / (32 * 28) = 73 -> 72 per thread; K1 uses 70).
#ifndef LZB_MAX_WARPS
#define LZB_MAX_WARPS 28
#endif

// Latency kernels (LAT): at most this many streams per SM; beyond it the throughput kernels win (measured crossover).
#ifndef LZB_LAT_WARPS
#define LZB_LAT_WARPS 8
#endif
// their shared-memory u16 per warp: small tables + the whole literal table in the reference's layout
static inline uint32_t lzb_lat_table_u16(uint32_t lclp) { return (uint32_t)T_LIT + (0x300u << lclp); }

struct LzbKC {
    uint32_t two, four, m1, m2017, k2048, shr11;
};
#define LZB_KC_INIT {2u, 4u, 0xFFFFFFFFu, (uint32_t)-2017, 2048u, 1u << 21}

// Shared-memory u16 per warp: small tables + plain literal columns.  Matched-literal columns (only touched by
// the first literal after a match, until its first mismatching bit) go to global memory (L1/L2): that halves
// the shared-memory footprint and raises residency from 14 to 28 streams per SM at lc+lp = 3.
static inline uint32_t lzb_table_u16(uint32_t lclp) { return (uint32_t)T_LIT + (0x100u << lclp); }
static inline uint32_t lzb_matched_u16(uint32_t lclp) { return 0x200u << lclp; }
static inline uint64_t lzb_carry_bytes(uint32_t lclp_cap) {
    return sizeof(LzbCarry) + (uint64_t)T_LIT * 2 + ((uint64_t)0x300 << lclp_cap) * 2;
}
