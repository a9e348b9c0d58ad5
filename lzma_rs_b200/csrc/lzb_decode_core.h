// lzb_decode_core.h -- K1's per-stream decode (one warp = one stream), written so that the SAME source also
// compiles as plain C++ with LZB_LANES == 1 (tests/host_emulation: CPU-side check of the decode logic against
// the oracle; test infrastructure only -- the shipped library always runs this code on the GPU).
//
// Behavioural contract: SURVEY.md 3.5 (bit-exact with the reference including its leniencies and its error
// precedence).  Every lane of a warp runs the same scalar decode (uniform control flow, shared-memory reads
// are broadcasts); lanes only differ inside the window / stored-chunk copies.
#pragma once
#include <stdint.h>

#include "lzb_types.h"

#ifdef __CUDACC__
#define LZB_LANES 32
#define LZB_DEV __device__ __forceinline__
#define LZB_DEV_NOINLINE __device__ __forceinline__
#define LZB_MEM __device__ __forceinline__
#define LZB_SYNCWARP() __syncwarp()
#define LZB_LDG(p) __ldg(p)
#define LZB_MIN(a, b) min(a, b)
#else  // host emulation of a 1-lane "warp"
#define LZB_LANES 1
#define LZB_DEV static inline
#define LZB_DEV_NOINLINE static
#define LZB_MEM inline
#define LZB_SYNCWARP() ((void)0)
#define LZB_LDG(p) (*(p))
#define LZB_MIN(a, b) ((a) < (b) ? (a) : (b))
#define __restrict__
static inline uint32_t __byte_perm(uint32_t x, uint32_t, uint32_t) { return __builtin_bswap32(x); }  // only 0x0123 is used
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { return (hi << s) | (lo >> (32 - s)); }
#endif

// Round-2 restructurings of the non-literal half of K1 (each can be switched off for an A/B measurement with
// tools/kbench.py: -DLZB_R2_xxx=0).  DESIGN.md section 4 has the measured effect of each.
#ifndef LZB_R2_DIRECT
#define LZB_R2_DIRECT 1  // rc_direct: carry-chain bit accumulation + min() code update, sentinel-terminated loop
#endif
#ifndef LZB_R2_TREES
#define LZB_R2_TREES 1   // align (and, == 1, the per-slot reverse trees) unrolled: 15 instead of 22 instructions per level
#endif
#ifndef LZB_R2_COPY
#define LZB_R2_COPY 1    // short non-overlapping matches: one predicated load/store pass, prev/match byte by shuffle
#endif
#ifndef LZB_R2_CHECKS
#define LZB_R2_CHECKS 1  // per-symbol limit checks folded into one compare pair, the ordered checks only behind it
#endif
#ifndef LZB_R2_STATE
#define LZB_R2_STATE 1   // arithmetic state transitions instead of the packed-nibble table; 3-instruction literal row
#endif

#define LZB_PRAGMA_(x) _Pragma(#x)
#define LZB_PRAGMA(x) LZB_PRAGMA_(x)
#define RC_TOP (1u << 24)
#define LZB_UNLIKELY(x) __builtin_expect(!!(x), 0)

// Keeps a loop-invariant value in a register instead of letting the compiler rematerialise it every symbol.
#ifdef __CUDACC__
#define LZB_KEEP(x) asm volatile("" : "+r"(x))
#else
#define LZB_KEEP(x) ((void)0)
#endif

// ------------------------------------------------------------------------------------------------
// range decoder + input window
// ------------------------------------------------------------------------------------------------
struct Dec {
    uint32_t range, code;
    const uint32_t* __restrict__ words;  // 4-byte aligned base of this stream's input
    uint32_t p;      // next byte to consume (relative to words)
    uint32_t lim;    // end of the current io::Take (lzma2.rs:189) or of the stream
    uint32_t lastw;  // last word index that may be loaded
    uint32_t cur;    // upcoming bytes, most significant first
    uint32_t nxt;    // prefetched following word (memory order)
};

LZB_DEV uint32_t ld_word(const Dec& d, uint32_t w) { return LZB_LDG(d.words + LZB_MIN(w, d.lastw)); }

LZB_DEV void rd_seek(Dec& d, uint32_t p) {
    uint32_t w = p >> 2;
    d.p = p;
    d.cur = __byte_perm(ld_word(d, w), 0, 0x0123) << ((p & 3u) * 8u);
    d.nxt = ld_word(d, w + 1);
}

// normalize, rangecoder.rs:60-69: ONE conditional 8-bit shift after every decision.  Reads past
// `lim` are detected by the caller as p > lim at the symbol boundary (UnexpectedEof, rangecoder.rs:64).
LZB_DEV void rc_normalize(Dec& d) {
    if (__builtin_expect(d.range < RC_TOP, 0)) {  // ~1 decision in 9: keep the common path fall-through
        d.range <<= 8;
        d.code = __funnelshift_l(d.cur, d.code, 8);  // (code << 8) | next byte
        d.cur <<= 8;
        d.p += 1;
        if (LZB_UNLIKELY((d.p & 3u) == 0)) {
            d.cur = __byte_perm(d.nxt, 0, 0x0123);
            d.nxt = ld_word(d, (d.p >> 2) + 1);
        }
    }
}

// The arithmetic of decode_bit (rangecoder.rs:93-120) on a probability already in a register.  Returns the bit
// (0/1); `np` = updated probability: one: p -= p >> 5; zero: p += (2048 - p) >> 5  ==  p + ((K - p) >>arith 5),
// K = one ? 31 : 2048 = 2048 - 2017 * bit  (floor((31 - p) / 32) == -(p >> 5)).  Normalisation is the caller's
// next step.  On the device the compare / range select / conditional code update are pinned in PTX so that the
// bit stays an opaque register for the constant-bank IMADs around it.
LZB_DEV uint32_t rc_step(Dec& d, const LzbKC& kc, uint32_t pv, uint32_t& np) {
    uint32_t bit, t;
#ifdef __CUDACC__
    asm("{\n\t.reg .pred p;\n\t.reg .u32 rs, bound, rb, kk;\n\t"
        "shr.u32 rs, %2, 11;\n\t"
        "mul.lo.u32 bound, rs, %4;\n\t"
        "setp.ge.u32 p, %3, bound;\n\t"
        "sub.u32 rb, %2, bound;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "selp.u32 %2, rb, bound, p;\n\t"
        "@p sub.u32 %3, %3, bound;\n\t"
        "mad.lo.u32 kk, %4, %5, 2048;\n\t"      // 2048 - p ...
        "@p add.u32 kk, kk, -2017;\n\t"         // ... or 31 - p: no register spent on the constant 31
        "mov.u32 %1, kk;\n\t}"
        : "=r"(bit), "=r"(t), "+r"(d.range), "+r"(d.code)
        : "r"(pv), "r"(kc.m1));
#else
    const uint32_t bound = (d.range >> 11) * pv;
    bit = d.code >= bound ? 1u : 0u;
    d.range = bit ? d.range - bound : bound;
    if (bit) d.code -= bound;
    t = pv * kc.m1 + (bit ? 31u : 2048u);  // K - p
#endif
    np = pv + (uint32_t)((int32_t)t >> 5);
    return bit;
}

// rc_step fused with the tree-node update.  A node is a handle-specific integer (TabSm: the byte address of the
// node's probability; TabPtr: its index); the child for `bit` is 2 * node + (bit ? x1 : x0) with per-tree constants
// x0 / x1 (TabSm: 0 - base and 2 - base; TabPtr: 0 and 1), so one select and one multiply-add replace
// "bit to register, m = 2m + bit, address = base + 2m".  The probability addend K = bit ? 31 : 2048 is selected
// from the same predicate.  Returns the child node; `np` = updated probability of the current node.
LZB_DEV uint32_t rc_step_tree(Dec& d, const LzbKC& kc, uint32_t pv, uint32_t& np, uint32_t node, uint32_t x0,
                              uint32_t x1) {
    uint32_t child, t;
#ifdef __CUDACC__
    // the multiply-adds with constant-bank multipliers run on the FMA pipe; shift, compare and three selects are
    // what is left for the ALU pipe (range >> 11 as mul.hi by 2^21 was measured slower: IMAD.HI is not full rate)
    asm("{\n\t.reg .pred p;\n\t.reg .u32 rs, bound, rb, kk, xx;\n\t"
        "shr.u32 rs, %2, 11;\n\t"
        "mul.lo.u32 bound, rs, %4;\n\t"
        "setp.ge.u32 p, %3, bound;\n\t"
        "sub.u32 rb, %2, bound;\n\t"
        "selp.u32 xx, %7, %6, p;\n\t"
        "selp.u32 %2, rb, bound, p;\n\t"
        "@p sub.u32 %3, %3, bound;\n\t"
        "mad.lo.u32 %0, %5, %8, xx;\n\t"
        "mad.lo.u32 kk, %4, %9, 2048;\n\t"
        "@p add.u32 kk, kk, -2017;\n\t"
        "mov.u32 %1, kk;\n\t}"
        : "=r"(child), "=r"(t), "+r"(d.range), "+r"(d.code)
        : "r"(pv), "r"(node), "r"(x0), "r"(x1), "r"(kc.two), "r"(kc.m1));
#else
    const uint32_t bound = (d.range >> 11) * pv;
    const uint32_t bit = d.code >= bound ? 1u : 0u;
    d.range = bit ? d.range - bound : bound;
    if (bit) d.code -= bound;
    child = node * kc.two + (bit ? x1 : x0);
    t = pv * kc.m1 + (bit ? 31u : 2048u);
#endif
    np = pv + (uint32_t)((int32_t)t >> 5);
    return child;
}

// ---- probability-table handles: u16 index in, value out -------------------------------------------------------
// TabPtr : plain pointers (host emulation; the global literal workspace of the lc+lp > 4 variant).
// TabSm  : 32-bit shared-space addresses with explicit ld/st.shared (device): addresses are IMADs off the
//          constant bank, and the compiler cannot turn the accesses into generic loads.
struct TabPtr {
    uint16_t* b;
    LZB_MEM uint32_t ld16(const LzbKC&, uint32_t i) const { return b[i]; }
    LZB_MEM void st16(const LzbKC&, uint32_t i, uint32_t v) const { b[i] = (uint16_t)v; }
    // tree-node interface (node = index)
    LZB_MEM uint32_t root(const LzbKC&) const { return 1u; }
    LZB_MEM uint32_t x0(const LzbKC&) const { return 0u; }
    LZB_MEM uint32_t x1(const LzbKC&) const { return 1u; }
    LZB_MEM uint32_t ldn(const LzbKC&, uint32_t node) const { return b[node]; }
    LZB_MEM void stn(const LzbKC&, uint32_t node, uint32_t v) const { b[node] = (uint16_t)v; }
    LZB_MEM uint32_t node_index(const LzbKC&, uint32_t node) const { return node; }
    // entries [i, i+1] (i even) / [i, i+3] (i a multiple of 4) as one / two 32-bit words, low entry in the low half
    LZB_MEM uint32_t ldw(const LzbKC&, uint32_t i) const { return (uint32_t)b[i] | ((uint32_t)b[i + 1] << 16); }
    LZB_MEM void ldq(const LzbKC& kc, uint32_t i, uint32_t& lo, uint32_t& hi) const {
        lo = ldw(kc, i);
        hi = ldw(kc, i + 2);
    }
    LZB_MEM TabPtr at(const LzbKC&, uint32_t off) const {
        TabPtr t = {b + off};
        return t;
    }
    LZB_MEM TabPtr at_bytes(uint32_t boff) const {
        TabPtr t = {b + (boff >> 1)};
        return t;
    }
};
#ifdef __CUDACC__
struct TabSm {
    uint32_t a;  // byte address in the shared window
    LZB_MEM uint32_t ld16(const LzbKC& kc, uint32_t i) const {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(i * kc.two + a) : "memory");
        return v;
    }
    LZB_MEM void st16(const LzbKC& kc, uint32_t i, uint32_t v) const {
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(i * kc.two + a), "h"((uint16_t)v) : "memory");
    }
    // tree-node interface (node = byte address of the node's probability)
    LZB_MEM uint32_t root(const LzbKC&) const { return a + 2u; }
    LZB_MEM uint32_t x0(const LzbKC&) const { return 0u - a; }
    LZB_MEM uint32_t x1(const LzbKC&) const { return 2u - a; }
    LZB_MEM uint32_t ldn(const LzbKC&, uint32_t node) const {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(node) : "memory");
        return v;
    }
    LZB_MEM void stn(const LzbKC&, uint32_t node, uint32_t v) const {
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(node), "h"((uint16_t)v) : "memory");
    }
    LZB_MEM uint32_t node_index(const LzbKC&, uint32_t node) const { return (node - a) >> 1; }
    // (latency kernels) entries [i, i+1] / [i, i+3]: the table is 4-byte / 8-byte aligned at entry i
    LZB_MEM uint32_t ldw(const LzbKC& kc, uint32_t i) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(i * kc.two + a) : "memory");
        return v;
    }
    LZB_MEM void ldq(const LzbKC& kc, uint32_t i, uint32_t& lo, uint32_t& hi) const {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(i * kc.two + a) : "memory");
    }
    LZB_MEM TabSm at(const LzbKC& kc, uint32_t off) const {
        TabSm t = {off * kc.two + a};
        return t;
    }
    LZB_MEM TabSm at_bytes(uint32_t boff) const {
        TabSm t = {boff + a};
        return t;
    }
};
#endif

// decode_bit on a table entry
template <class Tab>
LZB_DEV uint32_t rc_bit(Dec& d, const LzbKC& kc, const Tab& t, uint32_t idx) {
    uint32_t np;
    const uint32_t bit = rc_step(d, kc, t.ld16(kc, idx), np);
    t.st16(kc, idx, np);
    rc_normalize(d);
    return bit;
}

// get(count), rangecoder.rs:72-90 (count >= 1 at the only call site).  (Taking the bits in normalisation-free groups
// of 31 - clz(range) - 23 halvings was tried: same result, no gain.)
// Device form: the compare is the carry of code + ~range + 1 (set when code >= range: that is how ptxas lowers
// sub.cc, checked in the SASS: IADD3 ..., P0 / IMAD.X ..., P0); `addc` shifts it into the accumulator behind a
// sentinel 1, so the loop ends when the sentinel reaches bit `count` (no counter), and the conditional subtraction is
// min(code, code - range): when code < range the difference wraps above code.
// 10 instructions per bit instead of 13 (profiles/r01_k1_final.txt lines 1427-1437).
LZB_DEV uint32_t rc_direct(Dec& d, uint32_t count) {
#if defined(__CUDACC__) && LZB_R2_DIRECT
    uint32_t r = 1;
    const uint32_t end = 1u << count;
#pragma unroll 1
    do {
        d.range >>= 1;
        asm("{\n\t.reg .u32 t;\n\t"
            "sub.cc.u32 t, %1, %2;\n\t"   // carry = (code >= range) = the bit
            "addc.u32 %0, %0, %0;\n\t"    // r = 2r + bit
            "min.u32 %1, %1, t;\n\t}"
            : "+r"(r), "+r"(d.code)
            : "r"(d.range));
        rc_normalize(d);
    } while (r < end);
    return r - end;
#else
    uint32_t r = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < count; i++) {
        d.range >>= 1;
        const bool b = d.code >= d.range;
        if (b) d.code -= d.range;
        rc_normalize(d);
        r = (r << 1) | (b ? 1u : 0u);
    }
    return r;
#endif
}

// Bit-tree walk (parse_bit_tree, rangecoder.rs:122-134) of `nb` levels; `t` is the tree's table (node m at index m).
// Returns the final node index m in [2^nb, 2^(nb+1)): forward value = m - 2^nb; the reverse trees
// (parse_reverse_bit_tree, 136-151) walk the same nodes, so their value is the bit reversal of that.
template <bool UNROLL, int NB_CONST, class Tab>
LZB_DEV uint32_t rc_tree_walk(Dec& d, const LzbKC& kc, const Tab& t, uint32_t nb_rt) {
    const uint32_t nb = UNROLL ? (uint32_t)NB_CONST : nb_rt;
    const uint32_t x0 = t.x0(kc), x1 = t.x1(kc);
    uint32_t node = t.root(kc);
    uint32_t pv = t.ldn(kc, node);
#define LZB_TREE_STEP(I, NBV)                                                   \
    {                                                                           \
        uint32_t np;                                                            \
        const uint32_t child = rc_step_tree(d, kc, pv, np, node, x0, x1);       \
        t.stn(kc, node, np);                                                    \
        node = child;                                                           \
        if ((uint32_t)(I) + 1 < (NBV)) pv = t.ldn(kc, node);                    \
        rc_normalize(d);                                                        \
    }
    if (UNROLL) {
#pragma unroll
        for (int i = 0; i < NB_CONST; i++) LZB_TREE_STEP(i, (uint32_t)NB_CONST)
    } else {
#pragma unroll 1
        for (uint32_t i = 0; i < nb; i++) LZB_TREE_STEP(i, nb)
    }
#undef LZB_TREE_STEP
    return t.node_index(kc, node);
}

template <class Tab>
LZB_DEV uint32_t rc_tree_rt(Dec& d, const LzbKC& kc, const Tab& t, uint32_t nb) {
    return rc_tree_walk<false, 0>(d, kc, t, nb);
}

// Walk of nb <= MAXNB levels (nb >= 1), unrolled with an exit test per level: no loop counter, no constant reloads
// inside a rolled body (the rolled loop costs 22 instructions per level, an unrolled level 15 + the exit test).
template <int MAXNB, class Tab>
LZB_DEV uint32_t rc_tree_upto(Dec& d, const LzbKC& kc, const Tab& t, uint32_t nb) {
    const uint32_t x0 = t.x0(kc), x1 = t.x1(kc);
    uint32_t node = t.root(kc);
    uint32_t pv = t.ldn(kc, node);
#pragma unroll
    for (int i = 0; i < MAXNB; i++) {
        uint32_t np;
        const uint32_t child = rc_step_tree(d, kc, pv, np, node, x0, x1);
        t.stn(kc, node, np);
        node = child;
        const bool more = (uint32_t)(i + 1) < nb;
        if (i + 1 < MAXNB && more) pv = t.ldn(kc, node);
        rc_normalize(d);
        if (!more) break;
    }
    return t.node_index(kc, node);
}

// ---- latency form of the tree walks (LAT instantiations of K1: calls with so few streams that a warp has an SM
// sub-partition almost to itself, so the time of a decision is its dependent chain, not its instruction count) --------
// In the throughput kernels every level of a walk is load -> multiply -> compare -> child address -> load: ~55 cycles of
// which 29 are the shared-memory load.  Here the probabilities are fetched AHEAD of the decisions that select them: the
// children of node m are the adjacent entries (2m, 2m+1) and its grandchildren the four entries 4m .. 4m+3, so one 8-byte
// load issued when m becomes known delivers both candidates for the pair (= children) of the NEXT node; the chain per
// level is then select -> shift -> multiply -> compare (~20 cycles) and no load sits on it.
struct LatPre {  // entries 0 .. 7 of a tree: the root (1), its children (2, 3) and grandchildren (4 .. 7)
    uint32_t w0, w1, w2, w3;
};
template <bool A8, class Tab>
LZB_DEV LatPre lat_preload(const LzbKC& kc, const Tab& t) {
    LatPre p;
    if (A8) {  // the tree starts on an 8-byte boundary
        t.ldq(kc, 0, p.w0, p.w1);
        t.ldq(kc, 4, p.w2, p.w3);
    } else {
        p.w0 = t.ldw(kc, 0);
        p.w1 = t.ldw(kc, 2);
        p.w2 = t.ldw(kc, 4);
        p.w3 = t.ldw(kc, 6);
    }
    return p;
}
template <bool A8, class Tab>
LZB_DEV void lat_ldq(const LzbKC& kc, const Tab& t, uint32_t i, uint32_t& lo, uint32_t& hi) {
    if (A8) {
        t.ldq(kc, i, lo, hi);
    } else {
        lo = t.ldw(kc, i);
        hi = t.ldw(kc, i + 2);
    }
}

// State of a walk at node m: pv = p(m); (qp_lo, qp_hi) = the four entries below m's PARENT and pbit the decision taken
// there, i.e. m's children are the half selected by pbit; (qc_lo, qc_hi) = the four entries 4m .. 4m+3 (in flight).
// Every level issues the load for the node it arrives at and consumes the one issued two decisions earlier.
// `nb` levels (1 .. MAXNB) are walked; EXACT: nb == MAXNB.  `lim`: entries of the tree (no load reaches beyond it).
template <int MAXNB, bool A8, bool EXACT, class Tab>
LZB_DEV uint32_t lat_walk(Dec& d, const LzbKC& kc, const Tab& t, uint32_t m, uint32_t pv, uint32_t qp_lo, uint32_t qp_hi,
                          uint32_t pbit, uint32_t qc_lo, uint32_t qc_hi, uint32_t nb, uint32_t lim) {
    // (Tried: the normalisation bodies moved out of the fall-through path with gotos, so that the common case takes no
    // branch -- a lone warp pays ~20 cycles per taken branch.  nvcc / ptxas lay the blocks out inline again.)
#pragma unroll
    for (int i = 0; i < MAXNB; i++) {
        uint32_t np;
        const uint32_t bit = rc_step(d, kc, pv, np);
        t.st16(kc, m, np);
        const uint32_t pair = pbit ? qp_hi : qp_lo;  // entries 2m, 2m+1
        pv = bit ? pair >> 16 : pair & 0xFFFFu;
        m = m * 2u + bit;
        qp_lo = qc_lo;
        qp_hi = qc_hi;
        pbit = bit;
        // entries 4m .. 4m+3 of the node just reached: the probabilities of the level after next
        if (i + 3 < MAXNB && (EXACT || 4u * m + 3u < lim)) lat_ldq<A8>(kc, t, 4u * m, qc_lo, qc_hi);
        rc_normalize(d);
        if (!EXACT && (uint32_t)(i + 1) >= nb) break;
    }
    return m;
}

// Walk from the root with the first three levels preloaded (lat_preload).
template <int MAXNB, bool A8, bool EXACT, class Tab>
LZB_DEV uint32_t lat_tree(Dec& d, const LzbKC& kc, const Tab& t, const LatPre& pre, uint32_t nb) {
    return lat_walk<MAXNB, A8, EXACT>(d, kc, t, 1u, pre.w0 >> 16, pre.w1, pre.w1, 0u, pre.w2, pre.w3, nb, 1u << MAXNB);
}

// Rest of a literal walk from node m (1 <= m < 0x100) of the plain tree `t` (after the first mismatching bit of a
// matched literal, lzma.rs:540-556): up to 7 levels, rolled.  One exposed load at the start, look-ahead after it.
template <class Tab>
LZB_DEV uint32_t lat_lit_rest(Dec& d, const LzbKC& kc, const Tab& t, uint32_t m) {
    if (m >= 0x100u) return m;
    uint32_t pv = t.ld16(kc, m), pair = 0, qc_lo = 0, qc_hi = 0, pbit = 0;
    if (m < 0x80u) pair = t.ldw(kc, 2u * m);
    if (m < 0x40u) t.ldq(kc, 4u * m, qc_lo, qc_hi);
    uint32_t qp_lo = pair, qp_hi = pair;
#pragma unroll 1
    do {
        uint32_t np;
        const uint32_t bit = rc_step(d, kc, pv, np);
        t.st16(kc, m, np);
        const uint32_t pr = pbit ? qp_hi : qp_lo;
        pv = bit ? pr >> 16 : pr & 0xFFFFu;
        m = m * 2u + bit;
        qp_lo = qc_lo;
        qp_hi = qc_hi;
        pbit = bit;
        if (m < 0x40u) t.ldq(kc, 4u * m, qc_lo, qc_hi);
        rc_normalize(d);
    } while (m < 0x100u);
    return m;
}

// decode_bit on a probability fetched earlier (its address did not depend on the decisions in between)
template <class Tab>
LZB_DEV uint32_t rc_bit_pre(Dec& d, const LzbKC& kc, const Tab& t, uint32_t idx, uint32_t pv) {
    uint32_t np;
    const uint32_t bit = rc_step(d, kc, pv, np);
    t.st16(kc, idx, np);
    rc_normalize(d);
    return bit;
}

LZB_DEV uint32_t rev_bits(uint32_t v, uint32_t nb) {  // the low nb bits of v, reversed
#ifdef __CUDACC__
    return __brev(v) >> (32 - nb);
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < nb; i++) r |= ((v >> i) & 1u) << (nb - 1 - i);
    return r;
#endif
}

// Every probability = 0x400 (lzma.rs:188-214).  T is 16-byte aligned and n_u16 a multiple of 8 (all table sizes are).
// One rolled loop of 16-byte stores: this runs once per stream / state reset, and K1's code must stay inside the
// 32 KB instruction cache (an unrolled fill at four call sites cost 178 instructions).
LZB_DEV void fill_tables(uint16_t* T, uint32_t n_u16, int lane) {
#ifdef __CUDACC__
    uint4* T4 = reinterpret_cast<uint4*>(T);
    const uint4 v = make_uint4(0x04000400u, 0x04000400u, 0x04000400u, 0x04000400u);
#ifndef LZB_FILL_UNROLL
#define LZB_FILL_UNROLL 2
#endif
    LZB_PRAGMA(unroll LZB_FILL_UNROLL)
    for (uint32_t i = lane; i < n_u16 / 8; i += LZB_LANES) T4[i] = v;
#else
    (void)lane;
    for (uint32_t i = 0; i < n_u16; i++) T[i] = 0x400;
#endif
    LZB_SYNCWARP();
}

// ------------------------------------------------------------------------------------------------
// K1: decode one stream with one warp
// ------------------------------------------------------------------------------------------------
// The status and its message arguments are written behind opaque (volatile) moves: plain assignments of constants get
// hoisted ABOVE the branch into the hot path by the compiler (speculation is free in its cost model, but K1 is bound by
// instruction issue: 4 instructions per literal and per match were spent preparing errors that never happen, 1.8 % of
// all executed instructions in profiles/r01_k1_final.txt).
// Only in the instantiations without the host mirror: with it the same change measured 0.4 % slower (code layout), so
// those keep the plain form (FAIL is only used inside decode_item, where MIRROR is a template parameter).
#if defined(__CUDACC__) && !defined(LZB_FAIL_PLAIN)
#define FAIL(c, x, y)                                                                      \
    do {                                                                                   \
        if (MIRROR) {                                                                      \
            err = (c);                                                                     \
            ea0 = (uint64_t)(x);                                                           \
            ea1 = (uint64_t)(y);                                                           \
        } else {                                                                           \
            asm volatile("mov.u32 %0, %1;" : "=r"(err) : "r"((int32_t)(c)));               \
            asm volatile("mov.u64 %0, %1;" : "=l"(ea0) : "l"((uint64_t)(x)));              \
            asm volatile("mov.u64 %0, %1;" : "=l"(ea1) : "l"((uint64_t)(y)));              \
        }                                                                                  \
        goto finish;                                                                       \
    } while (0)
#else
#define FAIL(c, x, y)        \
    do {                     \
        err = (c);           \
        ea0 = (uint64_t)(x); \
        ea1 = (uint64_t)(y); \
        goto finish;         \
    } while (0)
#endif

// End of a literal (lzma.rs:299-307): limit checks, the byte, the state transition; STATE_EXPR = the new state.
#if LZB_R2_CHECKS
#define LZB_LIT_CHECKS                                                              \
    if (LZB_UNLIKELY((d.p > d.lim) | (opos >= lit_limit))) {                        \
        if (d.p > d.lim) FAIL(LZB_E_IO_EOF, 0, 0);                                  \
        if (opos >= mem_stop) FAIL(LZB_E_MEMLIMIT, itp->memlimit, 0);               \
        FAIL(LZB_E_CAPACITY, (uint64_t)opos + 1, 0);                                \
    }
#else
#define LZB_LIT_CHECKS                                                              \
    if (LZB_UNLIKELY(d.p > d.lim)) FAIL(LZB_E_IO_EOF, 0, 0);                        \
    if (LZB_UNLIKELY(opos >= lit_limit)) {                                          \
        if (LZB_UNLIKELY(opos >= mem_stop)) FAIL(LZB_E_MEMLIMIT, itp->memlimit, 0); \
        FAIL(LZB_E_CAPACITY, (uint64_t)opos + 1, 0);                                \
    }
#endif
#if LZB_R2_STATE
#define LZB_LIT_STATE(STATE_EXPR) state = (STATE_EXPR);
#else  // 0,0,0,0,1,2,3,4,5,6,4,5 as packed nibbles
#define LZB_LIT_STATE(STATE_EXPR) state = (uint32_t)(0x546543210000ull >> (state * 4)) & 0xFu;
#endif
#define LZB_LIT_TAIL(STATE_EXPR)                        \
    {                                                   \
        LZB_LIT_CHECKS                                  \
        prev_byte = sym & 0xFFu;                        \
        if (lane == 0) out[opos] = (uint8_t)prev_byte;  \
        opos += 1;                                      \
        LZB_LIT_STATE(STATE_EXPR)                       \
        mb_valid = false;                               \
        continue;                                       \
    }

// Warp copy of n bytes src -> dst (arbitrary, independent alignments; regions do not overlap): a few head bytes
// bring dst to a 4-byte boundary, the body stores aligned 32-bit words assembled from two aligned source words with
// a funnel shift (128 B per warp instruction instead of 32), the tail goes byte-wise.  SRC_CONST: the source is the
// read-only input blob (ld.global.nc).
// Register pressure matters beyond this function: inlined into decode_item with 4 vectors per lane in flight (LZB_COPY_U
// = 4, +20 live registers) it pushes ptxas into allocating K1's decode state in uniform registers, and the whole bit
// loop then runs on the (single) uniform datapath -- measured 44 % slower (profiles/r01_uniform_flip.md;
// tools/check_sass.py guards the build).  2 vectors per lane keep the normal allocation in every K1 variant; a
// __noinline__ call was tried and flips more variants, not fewer.  K4 / K6 (lzb_encode_kernels.cu) use it as is.
template <bool SRC_CONST>
LZB_DEV void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, int lane) {
#ifndef __CUDACC__
    (void)lane;
    for (uint32_t k = 0; k < n; k++) dst[k] = src[k];  // 1-lane emulation: plain copy
    return;
#endif
#ifdef __CUDACC__
    // dst is brought to a 16-byte boundary; the body stores aligned 16-byte vectors (512 B per warp instruction),
    // each assembled from five aligned source words with funnel shifts.  LZB_COPY_U vectors per lane are loaded before
    // the first is stored (SRC_CONST loads are non-coherent, so nothing orders them behind the stores).
    uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    if ((uint32_t)lane < head) dst[lane] = SRC_CONST ? LZB_LDG(src + lane) : src[lane];
    const uint32_t vecs = (n - head) >> 4;
    const uint8_t* s0 = src + head;
    const uint32_t sh = ((uint32_t)(uintptr_t)s0 & 3u) * 8u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s0 - ((uintptr_t)s0 & 3u));
    uint4* dv = reinterpret_cast<uint4*>(dst + head);
#ifndef LZB_COPY_U
#define LZB_COPY_U 2
#endif
    constexpr int U = LZB_COPY_U;
    for (uint32_t base = 0; base < vecs; base += U * LZB_LANES) {
        uint32_t w[U][5];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t k = base + u * LZB_LANES + lane;
            if (k < vecs) {
                const uint32_t* q = sw + 4u * k;
#pragma unroll
                for (int j = 0; j < 4; j++) w[u][j] = SRC_CONST ? __ldg(q + j) : q[j];
                // the fifth word is only dereferenced when the source is misaligned (it may lie past the last byte)
                w[u][4] = sh ? (SRC_CONST ? __ldg(q + 4) : q[4]) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t k = base + u * LZB_LANES + lane;
            if (k < vecs)
                dv[k] = make_uint4(__funnelshift_r(w[u][0], w[u][1], sh), __funnelshift_r(w[u][1], w[u][2], sh),
                                   __funnelshift_r(w[u][2], w[u][3], sh), __funnelshift_r(w[u][3], w[u][4], sh));
        }
    }
    const uint32_t done = head + vecs * 16u;
    if (done + (uint32_t)lane < n) dst[done + lane] = SRC_CONST ? LZB_LDG(src + done + lane) : src[done + lane];
#endif
}

// Warp fill of n bytes with one byte value (a dist == 1 match: the run-length case, BASELINE config 5).
LZB_DEV void warp_fill(uint8_t* dst, uint32_t byte, uint32_t n, int lane) {
#ifndef __CUDACC__
    (void)lane;
    for (uint32_t k = 0; k < n; k++) dst[k] = (uint8_t)byte;
    return;
#endif
    uint32_t head = (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u);
    if (head > n) head = n;
    if ((uint32_t)lane < head) dst[lane] = (uint8_t)byte;
    const uint32_t words = (n - head) >> 2;
    const uint32_t splat = byte * 0x01010101u;
#ifdef __CUDACC__
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
    for (uint32_t k = lane; k < words; k += LZB_LANES) dw[k] = splat;
#else
    (void)splat;
    for (uint32_t k = 0; k < words * 4; k++) dst[head + k] = (uint8_t)byte;
#endif
    const uint32_t done = head + words * 4u;
    if (done + (uint32_t)lane < n) dst[done + lane] = (uint8_t)byte;
}

// Copies out[from, to) (device window) to the same offsets of the host mirror; `from` is a multiple of 16 and both
// bases are 16-byte aligned (checked by the caller), so the body is 512-byte warp stores over PCIe.
LZB_DEV void mirror_to_host(const uint8_t* out, uint8_t* hout, uint32_t from, uint32_t to, int lane) {
    LZB_SYNCWARP();  // the window bytes were stored by other lanes
    // warp-uniform trip counts with predicated bodies: no lane-dependent loop exits inside the symbol loop
    const uint32_t to16 = from + ((to - from) & ~15u);
    for (uint32_t base = from; base < to16; base += 16u * LZB_LANES) {
        const uint32_t i = base + (uint32_t)lane * 16u;
        if (i < to16) {
#ifdef __CUDACC__
            *reinterpret_cast<uint4*>(hout + i) = *reinterpret_cast<const uint4*>(out + i);
#else
            for (int k = 0; k < 16; k++) hout[i + k] = out[i + k];
#endif
        }
    }
    if (to16 + (uint32_t)lane < to) hout[to16 + lane] = out[to16 + lane];  // < 16 trailing bytes
}

// The literal table (lzma.rs:194, [1 << (lc+lp)][0x300]) is split by column:
//   plain   columns 0x000..0x0FF  (every literal)                      -> `plain`,   row stride `plain_stride`
//   matched columns 0x100..0x2FF  (first literal after a match only)   -> `matched`, row stride `matched_stride`,
//                                                                         indexed (match_bit << 8) + node
// LIT_GLOBAL = false: plain columns in shared memory (T + T_LIT, stride 0x100), matched columns in the per-warp global
//                     workspace `gws` (stride 0x200); lc+lp <= tab_lclp <= 4, every LZMA2 stream.
// LIT_GLOBAL = true : .lzma streams with lc+lp > 4 (legal up to 12, lzma.rs:62-66): the whole table (up to 6 MiB) in
//                     `gws` with the reference's layout (stride 0x300; matched = +0x100).
// MainTab / PlainTab / MatchedTab: handle types of the small tables and the two literal parts.
// MIRROR: completed 4 KiB pages of the output are copied to the caller's pinned host buffer (itp->host_out) while
//         the stream is still decoding, so that the host API needs no device-to-host copy after the kernel.
// WIDE  : 1 = word-wide run fills, 2 = 16-byte vector stored-chunk copies.  Kept out of the default instantiation (0)
//         because K1 sits at the edge of the instruction cache: every extra path costs the common case 1-6 % even when
//         it never executes (measured); the host selects the variant per batch from the framing scan.
// CARRY : decompress::raw decoders -- the DecoderState a previous call left in an LzbCarry record (itp->host_out) is the
//         starting point, and what this call leaves is written back (lzb_types.h).  LIT_GLOBAL form only: the record's
//         literal area IS the kernel's literal workspace.
// LAT   : latency form for calls with few streams per SM (see lat_walk): the whole literal table in shared memory with
//         the reference's layout (plain = T + T_LIT, matched = + 0x100, stride 0x300, no global workspace), probabilities
//         fetched ahead of the decisions that select them.  More instructions per symbol, a much shorter dependent chain.
template <bool LIT_GLOBAL, bool MIRROR, int WIDE, bool CARRY = false, bool LAT = false, class MainTab, class PlainTab,
          class MatchedTab>
LZB_DEV_NOINLINE void decode_item(const LzbItem* __restrict__ itp, const uint8_t* __restrict__ in_blob,
                                  uint8_t* out_blob, uint16_t* T, uint16_t* gws, const MainTab tab,
                                  const PlainTab plain, const MatchedTab matched, const LzbKC kc_in, uint32_t tab_lclp,
                                  LzbResult* res, int lane) {
#if defined(__CUDACC__) && defined(LZB_R2_PINKC)
    // (experiment) multipliers pinned in registers instead of re-read from the constant bank at every use site
    LzbKC kcp = kc_in;
    kcp.two = kc_in.two + (tab_lclp >> 30);  // + 0, opaque to ptxas: a plain parameter load would be rematerialised
    kcp.m1 = kc_in.m1 - (tab_lclp >> 30);
    LZB_KEEP(kcp.two);
    LZB_KEEP(kcp.m1);
    const LzbKC& kc = kcp;
#else
    const LzbKC& kc = kc_in;
#endif
    Dec d;
    const bool is_lzma1 = itp->kind == LZB_ITEM_LZMA;
    const uint32_t p0 = (uint32_t)(itp->in_off & 3ull);
    const uint8_t* __restrict__ inb =
        ((itp->flags & LZB_ITEM_F_IN_FROM_OUT) ? (const uint8_t*)out_blob : in_blob) + (itp->in_off - p0);
    const uint32_t stream_lim = p0 + (uint32_t)itp->in_len;
    uint8_t* out = out_blob + itp->out_off;
    // (pinning this pointer in a register pair with an opaque asm was tried: ptxas then loses the address space (generic
    // ST instead of STG) and the warp-uniformity of everything derived from it -- 76 BSSY pairs in the bit loop)
    const uint32_t cap = (uint32_t)LZB_MIN(itp->out_cap, (uint64_t)0xFFFFF000u);
    const uint32_t tab_u16 = LIT_GLOBAL ? (uint32_t)T_LIT : T_LIT + ((LAT ? 0x300u : 0x100u) << tab_lclp);
    const uint32_t plain_stride = LIT_GLOBAL || LAT ? 0x300u : 0x100u, matched_stride = LIT_GLOBAL || LAT ? 0x300u : 0x200u;
    uint32_t opos = 0, dict_base = 0;
    uint32_t mirrored = 0;  // MIRROR: bytes already copied to the host buffer (multiple of 16)
    // first flush point staggered per stream: equal streams started together would otherwise all flush at the same
    // instant and serialise on the PCIe link (measured: +4.8 ms per 256 MiB, exactly one bulk copy)
    uint32_t mirror_next = 4096u + (uint32_t)((itp->out_off >> 4) * 2654435761ull >> 20 & 0xFF0u);
    uint8_t* const hout = MIRROR && !(CARRY && (itp->flags & LZB_ITEM_F_CARRY)) ? reinterpret_cast<uint8_t*>(itp->host_out) : nullptr;
    uint32_t state = 0, rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0;
    uint32_t lc = 0, lp = 0, pb = 0;
    uint32_t prev_byte = 0, match_byte = 0;
    bool mb_valid = false, tables_fresh = true;
    bool carry_live = false;  // CARRY: the tables in shared memory are this decoder's state (written back at the end)
    // (a run-time property of the work item inside the instantiations compiled with CARRY)
    LzbCarry* const carry = CARRY && (itp->flags & LZB_ITEM_F_CARRY) ? reinterpret_cast<LzbCarry*>(itp->host_out) : nullptr;
    uint32_t dict_size = 0xFFFFFFFFu, mem_stop = 0xFFFFFFFFu;
    uint32_t target = 0;
    bool has_target = false;
    int err = LZB_OK;
    uint64_t ea0 = 0, ea1 = 0;
    uint32_t chunks = 0;

    if (itp->kind == LZB_ITEM_PRESET) {  // status decided by the header parse (K2 / host)
        if (lane == 0) {
            res->code = itp->preset_code;
            res->chunks = 0;
            res->a0 = itp->preset_a0;
            res->a1 = 0;
            res->out_len = 0;
            res->sink_len = 0;
            res->consumed = 0;
        }
        return;
    }

    d.words = reinterpret_cast<const uint32_t*>(inb);
    d.lastw = itp->in_len ? (stream_lim - 1) >> 2 : 0;
    d.p = p0;
    d.lim = stream_lim;
    d.range = 0xFFFFFFFFu;
    d.code = 0;
    d.cur = d.nxt = 0;

    if (is_lzma1) {
        lc = itp->lc;
        lp = itp->lp;
        pb = itp->pb;
        dict_size = itp->dict_size;
        if (itp->memlimit < (uint64_t)dict_size) mem_stop = (uint32_t)itp->memlimit;  // lzbuffer.rs:209-217
        has_target = itp->unpacked != LZB_UNKNOWN_SIZE;
        target = (uint32_t)LZB_MIN(itp->unpacked, (uint64_t)0xFFFFFFFFu);  // sizes beyond the cap are never reached
        if (LZB_UNLIKELY(lc + lp > tab_lclp)) FAIL(LZB_E_UNSUPPORTED, lc + lp, tab_lclp);
    }
    if (CARRY && carry && !carry->fresh) {  // continue from the state the previous decompress() call left
        const uint16_t* saved = reinterpret_cast<const uint16_t*>(carry + 1);
        for (uint32_t i = lane; i < (uint32_t)T_LIT; i += LZB_LANES) T[i] = saved[i];
        LZB_SYNCWARP();
        state = carry->state;
        rep0 = carry->rep[0];
        rep1 = carry->rep[1];
        rep2 = carry->rep[2];
        rep3 = carry->rep[3];
        if (!is_lzma1) {
            lc = carry->lc;
            lp = carry->lp;
            pb = carry->pb;
        }
        tables_fresh = false;
    } else {
        fill_tables(T, tab_u16, lane);
        // global part: the whole literal table (LIT_GLOBAL; .lzma props never change mid-stream) or the matched columns
        if (!LAT) fill_tables(gws, LIT_GLOBAL ? 0x300u << (lc + lp) : 0x200u << tab_lclp, lane);
    }
    carry_live = CARRY && carry != nullptr;

    for (;;) {  // LZMA2 chunk loop (lzma2.rs:59-78); a .lzma stream is a single pass
        if (is_lzma1) {
            if (LZB_UNLIKELY(stream_lim - d.p < 5)) FAIL(LZB_E_LZMA_STREAM_TOO_SHORT, 0, 0);  // lzma.rs:643-644
        } else {
            if (LZB_UNLIKELY(d.p >= stream_lim)) FAIL(LZB_E_L2_STATUS_EOF, 0, 0);
            uint32_t status = inb[d.p];
            d.p += 1;
            chunks++;
            if (status == 0) break;
            if (status == 1 || status == 2) {  // parse_uncompressed, lzma2.rs:195-229
                if (LZB_UNLIKELY(stream_lim - d.p < 2)) FAIL(LZB_E_L2_UNPACKED_EOF, 0, 0);
                uint32_t n = (((uint32_t)inb[d.p] << 8) | inb[d.p + 1]) + 1;
                d.p += 2;
                if (status == 1) {  // accum.reset(): everything so far goes to the sink, window restarts
                    dict_base = opos;
                    prev_byte = 0;
                }
                if (LZB_UNLIKELY(stream_lim - d.p < n)) FAIL(LZB_E_L2_STORED_EOF, n, 0);
                if (LZB_UNLIKELY(cap - opos < n)) FAIL(LZB_E_CAPACITY, (uint64_t)opos + n, 0);
                if (WIDE == 2) {
                    warp_copy<true>(out + opos, inb + d.p, n, lane);
                } else {
                    for (uint32_t i = lane; i < n; i += LZB_LANES) out[opos + i] = LZB_LDG(inb + d.p + i);
                }
                opos += n;
                d.p += n;
                prev_byte = inb[d.p - 1];
                mb_valid = false;
                continue;
            }
            if (LZB_UNLIKELY(status < 0x80)) FAIL(LZB_E_L2_INVALID_STATUS, status, 0);  // lzma2.rs:94-99
            const uint32_t mode = (status >> 5) & 3u;
            if (LZB_UNLIKELY(stream_lim - d.p < 2)) FAIL(LZB_E_L2_UNPACKED_EOF, 0, 0);
            const uint32_t unpacked = ((((status & 0x1Fu) << 16) | ((uint32_t)inb[d.p] << 8) | inb[d.p + 1])) + 1;
            d.p += 2;
            if (LZB_UNLIKELY(stream_lim - d.p < 2)) FAIL(LZB_E_L2_PACKED_EOF, 0, 0);
            const uint32_t packed = (((uint32_t)inb[d.p] << 8) | inb[d.p + 1]) + 1;
            d.p += 2;
            if (mode == 3) {  // reset_dict
                dict_base = opos;
                prev_byte = 0;
            }
            if (mode >= 1) {      // reset_state
                if (mode >= 2) {  // reset_props
                    if (LZB_UNLIKELY(d.p >= stream_lim)) FAIL(LZB_E_L2_PROPS_EOF, 0, 0);
                    uint32_t props = inb[d.p];
                    d.p += 1;
                    if (LZB_UNLIKELY(props >= 225)) FAIL(LZB_E_L2_PROPS_RANGE, props, 0);
                    lc = props % 9;
                    props /= 9;
                    lp = props % 5;
                    pb = props / 5;
                    if (LZB_UNLIKELY(lc + lp > 4)) FAIL(LZB_E_L2_PROPS_LCLP, lc, lp);
                }
                if (LZB_UNLIKELY(lc + lp > tab_lclp)) FAIL(LZB_E_UNSUPPORTED, lc + lp, tab_lclp);
                if (!tables_fresh) {  // reset_state, lzma.rs:216-249
                    fill_tables(T, tab_u16, lane);
                    if (!LAT) fill_tables(gws, LIT_GLOBAL ? 0x300u << (lc + lp) : 0x200u << tab_lclp, lane);
                } else if (LIT_GLOBAL) {  // the table was initialised for the old lc+lp: cover the new one
                    fill_tables(gws, 0x300u << (lc + lp), lane);
                }
                state = 0;
                rep0 = rep1 = rep2 = rep3 = 0;
            }
            has_target = true;
            target = (opos - dict_base) + unpacked;  // set_unpacked_size(unpacked + accum.len()), lzma2.rs:186-187
            d.lim = LZB_MIN(d.p + packed, stream_lim);   // input.take(packed_size), lzma2.rs:189
            if (LZB_UNLIKELY(d.lim - d.p < 5)) FAIL(LZB_E_L2_INPUT_TOO_SHORT, 0, 0);
            mb_valid = false;
        }
        tables_fresh = false;

        // RangeDecoder::new, rangecoder.rs:20-30: one byte skipped (value ignored), then a BE u32
        d.range = 0xFFFFFFFFu;
        d.code = ((uint32_t)inb[d.p + 1] << 24) | ((uint32_t)inb[d.p + 2] << 16) | ((uint32_t)inb[d.p + 3] << 8) |
                 (uint32_t)inb[d.p + 4];
        rd_seek(d, d.p + 5);

        uint32_t pb_mask = (1u << pb) - 1, lp_mask = (1u << lp) - 1, lit_shift = 8 - lc;
        // symbol-loop exits as absolute output positions: `stop_at` = where the expected size is reached
        // (lzma.rs:442-445); `lit_limit` = first position a literal may not be written to (capacity / memlimit)
        uint32_t stop_at = has_target ? dict_base + target : 0xFFFFFFFFu;
        if (has_target && target > 0xFFFFFFFFu - dict_base) stop_at = 0xFFFFFFFFu;
        uint32_t lit_limit = cap < mem_stop ? cap : mem_stop;
#if LZB_R2_CHECKS
        // first position a match may not reach: capacity / memlimit, and for LZMA2 the chunk's declared size
        uint32_t match_limit = is_lzma1 ? lit_limit : LZB_MIN(lit_limit, stop_at);
        LZB_KEEP(match_limit);
#endif
#if LZB_R2_STATE
        uint32_t row_mask = (1u << (lc + lp)) - 1;
        LZB_KEEP(row_mask);
        uint32_t plain_row_bytes = plain_stride * 2u;  // pinned: row address = one multiply-add
        LZB_KEEP(plain_row_bytes);
#endif
        LZB_KEEP(pb_mask);
        LZB_KEEP(lp_mask);
        LZB_KEEP(lit_shift);
        LZB_KEEP(stop_at);
        LZB_KEEP(lit_limit);

        // process_mode(Finish), lzma.rs:435-455, 496-511
        for (;;) {
            // All lanes update the probability tables redundantly, which is only sound while they run in lockstep (every
            // lane has loaded a probability before any lane stores its update).  The throughput kernels are convergent by
            // construction (tools/check_sass.py guards it); the LIT_GLOBAL instantiations (lc+lp > 4 kernel, raw decoder
            // objects) are compiled with reconvergence barriers all over the bit loop, i.e. the lanes MAY drift apart after
            // a lane-dependent branch (the literal store of lane 0) -- observed on the device as nondeterministic decode
            // errors of the raw decoder kernel.  Re-align them once per symbol; these are latency paths.  The `fill`
            // instantiations carry the barriers too since the dist-1 shortcut (one WARPSYNC per 273-byte symbol there).
            if (LIT_GLOBAL || WIDE == 1 || LAT) LZB_SYNCWARP();
            if (MIRROR && LZB_UNLIKELY(opos >= mirror_next) && hout) {
                const uint32_t upto = opos & ~15u;
                mirror_to_host(out, hout, mirrored, upto, lane);
                mirrored = upto;
                mirror_next = upto + 4096u;
            }
            if (opos >= stop_at) break;
            if (LZB_UNLIKELY(d.code == 0) && !has_target && d.p == d.lim) break;  // is_finished_ok, rangecoder.rs:50-52
            const uint32_t len = opos - dict_base;
            const uint32_t pos_state = len & pb_mask;

            // literal context row (decode_literal, lzma.rs:526-538).  After a literal (state < 7) prev_byte is in a
            // register, so the root of the plain literal tree is fetched while is_match is being decoded.
#if LZB_R2_STATE
            // ((len & lp_mask) << lc) + (prev_byte >> (8 - lc)) as one shift of len:prev_byte
#ifdef __CUDACC__
            const uint32_t lit_row = (__byte_perm(prev_byte, len, 0x6540) >> lit_shift) & row_mask;  // len : prev_byte
#else
            const uint32_t lit_row = (((len << 8) | prev_byte) >> lit_shift) & row_mask;
#endif
#else
            const uint32_t lit_row = ((len & lp_mask) << lc) + (prev_byte >> lit_shift);
#endif
#if LZB_R2_STATE
            const PlainTab probs = plain.at_bytes(lit_row * plain_row_bytes);
#else
            const PlainTab probs = plain.at(kc, lit_row * plain_stride);
#endif
            const uint32_t i_is_match = T_IS_MATCH + (state << 4) + pos_state;
            const uint32_t p_is_match = tab.ld16(kc, i_is_match);
#if LZB_R2_STATE
            // (unconditional: after a match the value is simply not used -- cheaper than the predicate and the zero)
            const uint32_t lit_pv = LAT ? 0u : probs.ldn(kc, probs.root(kc));
#else
            uint32_t lit_pv = 0;
            if (state < 7) lit_pv = probs.ldn(kc, probs.root(kc));
#endif
            // LAT: the first three levels of the plain literal tree and is_rep's probability travel with is_match's
            LatPre lit_pre = {0u, 0u, 0u, 0u};
            uint32_t p_is_rep = 0;
            if (LAT) {
                lit_pre = lat_preload<true>(kc, probs);
                p_is_rep = tab.ld16(kc, T_IS_REP + state);
            }
            uint32_t np_im;
            const uint32_t is_lz = rc_step(d, kc, p_is_match, np_im);
            tab.st16(kc, i_is_match, np_im);
            rc_normalize(d);

            if (!is_lz) {
                // ---- literal, lzma.rs:287-307 + decode_literal 526-561
                uint32_t sym = 1;
                if (state >= 7) {
                    if (LZB_UNLIKELY(!mb_valid)) {  // last_n(rep[0] + 1), lzbuffer.rs:98-108 / 240-256
                        if (LZB_UNLIKELY(d.p > d.lim)) FAIL(LZB_E_IO_EOF, 0, 0);
                        if (LZB_UNLIKELY(rep0 >= dict_size)) FAIL(LZB_E_MATCH_DIST_DICT, (uint64_t)rep0 + 1, dict_size);
                        if (LZB_UNLIKELY(rep0 >= len)) FAIL(LZB_E_MATCH_DIST_OUT, (uint64_t)rep0 + 1, len);
                        LZB_SYNCWARP();
                        match_byte = out[opos - rep0 - 1];
                    }
                    uint32_t mb = match_byte;
                    const MatchedTab mprobs = matched.at(kc, lit_row * matched_stride);
                    if (LAT) {
                        // the eight probabilities along the path the match byte predicts are fetched at once: level k is
                        // entry (match_bit_k << 8) + sym_k, sym_k = the top k bits of the match byte behind a leading 1
                        const uint32_t mbx = 0x100u | match_byte;
                        uint32_t pk[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) pk[k] = mprobs.ld16(kc, (((mbx >> (7 - k)) & 1u) << 8) + (mbx >> (8 - k)));
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const uint32_t match_bit = (mbx >> (7 - k)) & 1u;
                            const uint32_t bit = rc_bit_pre(d, kc, mprobs, (match_bit << 8) + sym, pk[k]);
                            sym = (sym << 1) | bit;
                            if (match_bit != bit) break;
                        }
                        sym = lat_lit_rest(d, kc, probs, sym);  // the plain tree from the first mismatch on
                    } else {
#pragma unroll 1
                        do {
                            const uint32_t match_bit = (mb >> 7) & 1u;
                            mb <<= 1;
                            const uint32_t bit = rc_bit(d, kc, mprobs, (match_bit << 8) + sym);
                            sym = (sym << 1) | bit;
                            if (match_bit != bit) break;
                        } while (sym < 0x100);
#pragma unroll 1
                        while (sym < 0x100) sym = (sym << 1) | rc_bit(d, kc, probs, sym);
                    }
                } else if (LAT) {
                    sym = lat_tree<8, true, true>(d, kc, probs, lit_pre, 8);
                    LZB_LIT_TAIL(state > 3 ? state - 3 : 0u)
                } else {  // plain 8-level walk (root fetched above, while is_match was being decoded)
                    const uint32_t x0 = probs.x0(kc), x1 = probs.x1(kc);
                    uint32_t node = probs.root(kc), pv = lit_pv;
#ifndef LZB_LIT_UNROLL
#define LZB_LIT_UNROLL 8
#endif
                    LZB_PRAGMA(unroll LZB_LIT_UNROLL)
                    for (int i = 0; i < 8; i++) {
                        uint32_t np;
                        const uint32_t child = rc_step_tree(d, kc, pv, np, node, x0, x1);
                        probs.stn(kc, node, np);
                        node = child;
                        if (i < 7) pv = probs.ldn(kc, node);
                        rc_normalize(d);
                    }
                    sym = probs.node_index(kc, node);
#if LZB_R2_STATE
                    LZB_LIT_TAIL(state > 3 ? state - 3 : 0u)  // 0..3 -> 0; 4,5,6 -> 1,2,3  (own copy: no join with the matched path)
#endif
                }
                LZB_LIT_TAIL(state < 10 ? state - 3 : state - 6)  // 7,8,9 -> 4,5,6; 10,11 -> 4,5
            }

            // ---- LZ, lzma.rs:309-390
            uint32_t mlen;
            if (LAT) {
                // every probability whose address depends on state / pos_state only is fetched now and used up to five
                // decisions later; the length trees of this pos_state are fetched whole (8 entries each)
                const uint32_t p_g0 = tab.ld16(kc, T_IS_REP_G0 + state);
                const uint32_t p_r0l = tab.ld16(kc, T_IS_REP0LONG + (state << 4) + pos_state);
                const uint32_t p_g1 = tab.ld16(kc, T_IS_REP_G1 + state);
                const uint32_t p_g2 = tab.ld16(kc, T_IS_REP_G2 + state);
                const uint32_t ch_len = tab.ldw(kc, T_LEN), ch_rep = tab.ldw(kc, T_REP_LEN);  // choice : choice2
                const bool is_rep = rc_bit_pre(d, kc, tab, T_IS_REP + state, p_is_rep) != 0;
                const uint32_t L = is_rep ? (uint32_t)T_REP_LEN : (uint32_t)T_LEN;
                const uint32_t ch = is_rep ? ch_rep : ch_len;
                const MainTab t_low = tab.at(kc, L + T_LEN_LOW + pos_state * 8), t_mid = tab.at(kc, L + T_LEN_MID + pos_state * 8);
                const MainTab t_high = tab.at(kc, L + T_LEN_HIGH);
                const LatPre pre_low = lat_preload<false>(kc, t_low), pre_mid = lat_preload<false>(kc, t_mid);
                // the distance slot tree of lengths >= 5 (the usual case), ahead of the length decode
                LatPre pre_ps3 = {0u, 0u, 0u, 0u};
                if (!is_rep) pre_ps3 = lat_preload<true>(kc, tab.at(kc, T_POS_SLOT + 3 * 64));
                bool short_rep = false;
                if (is_rep) {  // lzma.rs:312-345
                    if (!rc_bit_pre(d, kc, tab, T_IS_REP_G0 + state, p_g0)) {
                        if (!rc_bit_pre(d, kc, tab, T_IS_REP0LONG + (state << 4) + pos_state, p_r0l)) short_rep = true;
                    } else {
                        uint32_t dist;
                        if (!rc_bit_pre(d, kc, tab, T_IS_REP_G1 + state, p_g1)) {
                            dist = rep1;
                        } else {
                            if (!rc_bit_pre(d, kc, tab, T_IS_REP_G2 + state, p_g2)) {
                                dist = rep2;
                            } else {
                                dist = rep3;
                                rep3 = rep2;
                            }
                            rep2 = rep1;
                        }
                        rep1 = rep0;
                        rep0 = dist;
                    }
                } else {  // lzma.rs:355-361
                    rep3 = rep2;
                    rep2 = rep1;
                    rep1 = rep0;
                }
                if (short_rep) {
                    state = state < 7 ? 9 : 11;
                    mlen = 1;
                } else {
                    // LenDecoder::decode, rangecoder.rs:256-269
                    uint32_t l;
                    const uint32_t c1 = rc_bit_pre(d, kc, tab, L + 0, ch & 0xFFFFu);
                    if (!c1) {
                        l = lat_tree<3, false, true>(d, kc, t_low, pre_low, 3) - 8;
                    } else {
                        const LatPre pre_high = lat_preload<false>(kc, t_high);
                        if (!rc_bit_pre(d, kc, tab, L + 1, ch >> 16))
                            l = lat_tree<3, false, true>(d, kc, t_mid, pre_mid, 3) - 8 + 8;
                        else
                            l = lat_tree<8, false, true>(d, kc, t_high, pre_high, 8) - 256 + 16;
                    }
                    if (is_rep) {
                        state = state < 7 ? 8 : 11;
                    } else {
                        state = state < 7 ? 7 : 10;
                        // decode_distance, lzma.rs:563-592
                        LatPre pre_ps = pre_ps3;
                        if (l < 3) pre_ps = lat_preload<true>(kc, tab.at(kc, T_POS_SLOT + l * 64));
                        const uint32_t pos_slot =
                            lat_tree<6, true, true>(d, kc, tab.at(kc, T_POS_SLOT + (l < 3 ? l : 3) * 64), pre_ps, 6) - 64;
                        if (pos_slot < 4) {
                            rep0 = pos_slot;
                        } else {
                            const uint32_t nd = (pos_slot >> 1) - 1;
                            uint32_t r = (2u | (pos_slot & 1u)) << nd;
                            if (pos_slot < 14) {
                                const uint32_t off = 2u * ((1u << nd) - 2u) + ((pos_slot & 1u) << nd);
                                const MainTab t_pd = tab.at(kc, T_POS_DEC + off);
                                const LatPre pre_pd = lat_preload<false>(kc, t_pd);
                                r += rev_bits(lat_tree<5, false, false>(d, kc, t_pd, pre_pd, nd), nd);
                            } else {
                                const MainTab t_al = tab.at(kc, T_ALIGN);
                                const LatPre pre_al = lat_preload<true>(kc, t_al);  // arrives while the direct bits are taken
                                r += rc_direct(d, nd - 4) << 4;
                                r += rev_bits(lat_tree<4, true, true>(d, kc, t_al, pre_al, 4), 4);
                            }
                            rep0 = r;
                        }
                        if (LZB_UNLIKELY(d.p > d.lim)) FAIL(LZB_E_IO_EOF, 0, 0);
                        if (rep0 == 0xFFFFFFFFu) {  // end-of-stream marker, lzma.rs:373-381
                            if (d.code == 0 && d.p == d.lim) goto chunk_done;
                            FAIL(LZB_E_EOS_MORE_BYTES, 0, 0);
                        }
                    }
                    mlen = l + 2;
                }
                if (LZB_UNLIKELY(d.p > d.lim)) FAIL(LZB_E_IO_EOF, 0, 0);
            } else {
                const bool is_rep = rc_bit(d, kc, tab, T_IS_REP + state) != 0;
                bool short_rep = false;
                if (is_rep) {  // lzma.rs:312-345
                    if (!rc_bit(d, kc, tab, T_IS_REP_G0 + state)) {
                        if (!rc_bit(d, kc, tab, T_IS_REP0LONG + (state << 4) + pos_state)) short_rep = true;
                    } else {
                        uint32_t dist;
                        if (!rc_bit(d, kc, tab, T_IS_REP_G1 + state)) {
                            dist = rep1;
                        } else {
                            if (!rc_bit(d, kc, tab, T_IS_REP_G2 + state)) {
                                dist = rep2;
                            } else {
                                dist = rep3;
                                rep3 = rep2;
                            }
                            rep2 = rep1;
                        }
                        rep1 = rep0;
                        rep0 = dist;
                    }
                } else {  // lzma.rs:355-361
                    rep3 = rep2;
                    rep2 = rep1;
                    rep1 = rep0;
                }
                if (short_rep) {
                    state = state < 7 ? 9 : 11;
                    mlen = 1;
                } else {
                    // LenDecoder::decode, rangecoder.rs:256-269 (len_decoder / rep_len_decoder share this code)
                    const uint32_t L = is_rep ? (uint32_t)T_REP_LEN : (uint32_t)T_LEN;
                    uint32_t l;
                    const uint32_t c1 = rc_bit(d, kc, tab, L + 0);
                    const uint32_t c2 = c1 ? rc_bit(d, kc, tab, L + 1) : 0u;
                    if (!c2) {  // low (c1 == 0) and mid (c1 == 1) coders: one 3-level walk site
                        const uint32_t base = L + (c1 ? (uint32_t)T_LEN_MID : (uint32_t)T_LEN_LOW) + pos_state * 8;
                        l = rc_tree_walk<true, 3>(d, kc, tab.at(kc, base), 3) - 8 + c1 * 8;
                    } else {
#if LZB_R2_TREES
                        if (WIDE == 1)  // run-length batches (config 5) take this coder on every symbol
                            l = rc_tree_walk<true, 8>(d, kc, tab.at(kc, L + T_LEN_HIGH), 8) - 256 + 16;
                        else
#endif
                            l = rc_tree_rt(d, kc, tab.at(kc, L + T_LEN_HIGH), 8) - 256 + 16;
                    }
                    if (is_rep) {
                        state = state < 7 ? 8 : 11;
                    } else {
                        state = state < 7 ? 7 : 10;
                        // decode_distance, lzma.rs:563-592
                        const uint32_t pos_slot = rc_tree_walk<true, 6>(d, kc, tab.at(kc, T_POS_SLOT + (l < 3 ? l : 3) * 64), 6) - 64;
                        if (pos_slot < 4) {
                            rep0 = pos_slot;
                        } else {
                            const uint32_t nd = (pos_slot >> 1) - 1;
                            uint32_t r = (2u | (pos_slot & 1u)) << nd;
                            if (pos_slot < 14) {  // per-slot reverse tree (own aligned block, see lzb_types.h)
                                const uint32_t off = 2u * ((1u << nd) - 2u) + ((pos_slot & 1u) << nd);
#if LZB_R2_TREES == 1
                                r += rev_bits(rc_tree_upto<5>(d, kc, tab.at(kc, T_POS_DEC + off), nd), nd);
#else
                                r += rev_bits(rc_tree_rt(d, kc, tab.at(kc, T_POS_DEC + off), nd), nd);
#endif
                            } else {
                                r += rc_direct(d, nd - 4) << 4;
#if LZB_R2_TREES
                                r += rev_bits(rc_tree_walk<true, 4>(d, kc, tab.at(kc, T_ALIGN), 4), 4);
#else
                                r += rev_bits(rc_tree_rt(d, kc, tab.at(kc, T_ALIGN), 4), 4);
#endif
                            }
                            rep0 = r;
                        }
                        if (LZB_UNLIKELY(d.p > d.lim)) FAIL(LZB_E_IO_EOF, 0, 0);
                        if (rep0 == 0xFFFFFFFFu) {  // end-of-stream marker, lzma.rs:373-381
                            if (d.code == 0 && d.p == d.lim) goto chunk_done;  // Finished: fall to the size check
                            FAIL(LZB_E_EOS_MORE_BYTES, 0, 0);
                        }
                    }
                    mlen = l + 2;
                }
                if (LZB_UNLIKELY(d.p > d.lim)) FAIL(LZB_E_IO_EOF, 0, 0);
            }

            // ---- append_lz(mlen, rep0 + 1), lzbuffer.rs:125-143 / 272-297
            {
#if LZB_R2_CHECKS
                // one compare pair on the common path; the reference's ordered checks only behind it (any failure of
                // the pair implies that one of them fails: match_limit is the smallest of their limits)
                if (LZB_UNLIKELY((rep0 >= LZB_MIN(dict_size, len)) | (match_limit - opos < mlen)))
#endif
                {
                    if (LZB_UNLIKELY(rep0 >= dict_size)) FAIL(LZB_E_LZ_DIST_DICT, (uint64_t)rep0 + 1, dict_size);
                    if (LZB_UNLIKELY(rep0 >= len)) FAIL(LZB_E_LZ_DIST_OUT, (uint64_t)rep0 + 1, len);
                    if (LZB_UNLIKELY(mem_stop - opos < mlen && mem_stop != 0xFFFFFFFFu)) FAIL(LZB_E_MEMLIMIT, itp->memlimit, 0);
                    if (!is_lzma1 && len + mlen > target)  // the window is never flushed past this point
                        FAIL(LZB_E_UNPACKED_MISMATCH, target, (uint64_t)len + mlen);
                    if (LZB_UNLIKELY(cap - opos < mlen)) FAIL(LZB_E_CAPACITY, (uint64_t)opos + mlen, 0);
                }
                const uint32_t dist = rep0 + 1;
                const uint8_t* src = out + opos - dist;
                uint8_t* dst = out + opos;
                LZB_SYNCWARP();  // earlier stores of other lanes are visible to these loads
                uint32_t i_last = mlen - 1, i_next = mlen;
#if defined(__CUDACC__) && LZB_R2_COPY
                if (mlen < 32u && dist >= mlen) {
                    // the common match: one predicated pass.  Lanes 0..mlen-1 move the bytes; lane mlen also loads the byte
                    // behind the source range (= out[new opos - dist], the match byte of a following literal) unless that is
                    // dst[0] itself (dist == mlen), which lane 0 holds.  prev_byte / match_byte then come from a shuffle
                    // instead of two more global round trips on the next symbol's dependency chain.
                    const uint32_t nload = mlen + (dist > mlen ? 1u : 0u);
                    uint32_t v = 0;
                    if ((uint32_t)lane < nload) v = src[lane];
                    if ((uint32_t)lane < mlen) dst[lane] = (uint8_t)v;
                    prev_byte = __shfl_sync(0xffffffffu, v, (int)i_last);
                    match_byte = __shfl_sync(0xffffffffu, v, (int)(dist > mlen ? mlen : 0u));
                } else
#endif
                if (dist >= mlen) {
                    for (uint32_t i = lane; i < mlen; i += LZB_LANES) dst[i] = src[i];
                    if (dist == mlen) i_next = 0;
                    prev_byte = src[i_last];   // last byte written      (uniform load)
                    match_byte = src[i_next];  // out[new_opos - dist]: the match byte of a following literal
                } else {  // overlapping: the window replicates with period dist
                    // (identical index expressions on both sides: written as src[0] the compiler folds the load into the
                    // lane-dependent fill and loses the warp-uniformity of prev_byte downstream)
#if LZB_R2_COPY
                    if (dist == 1) {
                        // run of the previous byte (zero fills in real data; every symbol of BASELINE config 5): that
                        // byte is already in a register -- it is the literal context -- so the next symbol's prev_byte /
                        // match_byte need no memory round trip and no index reduced modulo dist.  In the `fill`
                        // instantiation this costs ptxas its proof that prev_byte is warp-uniform (85 BSSY pairs in the
                        // bit loop, whichever way the fill value is produced) and is still 13 % faster on config 5
                        // (628 -> 710 GB/s): two global round trips per 273-byte symbol weigh more than the barriers.
                        match_byte = prev_byte;
                    } else
#endif
                    {
                        prev_byte = src[i_last % dist];
                        match_byte = src[i_next % dist];
                    }
                    if (WIDE == 1 && dist == 1) {  // run of one byte (BASELINE config 5): word-wide fill
                        warp_fill(dst, prev_byte, mlen, lane);
                    } else {
                        for (uint32_t i = lane; i < mlen; i += LZB_LANES) dst[i] = src[i % dist];
                    }
                }
                mb_valid = true;
                opos += mlen;
            }
        }
    chunk_done:
        {
            const uint32_t len = opos - dict_base;
            if (has_target && len != target)  // lzma.rs:513-521
                FAIL(LZB_E_UNPACKED_MISMATCH, is_lzma1 ? itp->unpacked : (uint64_t)target, len);
        }
        if (is_lzma1) break;
    }

finish:
    LZB_SYNCWARP();
    if (CARRY && carry_live) {  // what the next decompress() call of this decoder starts from
        uint16_t* saved = reinterpret_cast<uint16_t*>(carry + 1);
        for (uint32_t i = lane; i < (uint32_t)T_LIT; i += LZB_LANES) saved[i] = T[i];
        if (lane == 0) {
            carry->fresh = 0;
            carry->state = state;
            carry->rep[0] = rep0;
            carry->rep[1] = rep1;
            carry->rep[2] = rep2;
            carry->rep[3] = rep3;
            carry->lc = lc;
            carry->lp = lp;
            carry->pb = pb;
        }
    }
    if (MIRROR && hout && opos > mirrored) mirror_to_host(out, hout, mirrored, opos, lane);
    if (lane == 0) {
        uint64_t sink = opos;
        if (err != LZB_OK) {
            // what had reached the caller's sink when the reference returned the error:
            // LZMA2: everything before the last dict reset (lzbuffer.rs:73-78); LZMA: whole ring slabs (264-267)
            sink = is_lzma1 ? (uint64_t)(opos / dict_size) * dict_size : dict_base;
        }
        res->code = err;
        res->chunks = chunks;
        res->a0 = ea0;
        res->a1 = ea1;
        res->out_len = opos;
        res->sink_len = sink;
        res->consumed = d.p - p0;
    }
}

