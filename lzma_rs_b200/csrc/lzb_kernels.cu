// lzb_kernels.cu -- sm_100a kernels of the many-stream LZMA / LZMA2 / XZ decoder.
//
//   K1  lzb_decode_kernel   one warp = one independent stream: LZMA2 chunk walk (lzma2.rs:52-229),
//                           range decoder (rangecoder.rs:60-151), symbol state machine
//                           (lzma.rs:278-393, 526-592) and LZ window copy (lzbuffer.rs:125-143,272-297)
//                           fused in one persistent kernel; probability tables live in shared memory.
//   K2  lzb_scan_kernel     one thread = one stream: .lzma header parse (lzma.rs:96-161) or LZMA2
//                           chunk-header walk (lzma2.rs:128-136,204-207) -> work items + size summary.
//   K3  lzb_crc_*           CRC-32/ISO-HDLC + CRC-64/XZ of decoded ranges (xz.rs:295-333).
//
// Behavioural contract: SURVEY.md 3.5 (bit-exact with the reference including its leniencies and
// its error precedence).  Every lane of a warp runs the same scalar decode (uniform control flow,
// shared-memory reads are broadcasts); lanes only differ inside the window/stored copies.
#include <cuda_runtime.h>
#include <stdint.h>

#include "lzb_types.h"

#define FULL_MASK 0xffffffffu

#include "lzb_decode_core.h"

// ------------------------------------------------------------------------------------------------
// K1 launcher
// ------------------------------------------------------------------------------------------------
// Persistent kernels: warps pull stream indices (pre-sorted longest first by the host) from a counter.
// SCHED: the launch carries a placement plan (lzb_sched.h): order[0 .. n_static) holds the first item of every warp
// (n_static = gridDim.x * warps), LZB_ORDER_PARK there = this warp stays out of the launch; order[n_static ..) is the
// dynamic queue.  A separate instantiation because K1's hot loop is sensitive to code layout: the same loop head in
// the default kernels costs batches that need no plan 0.5-3 % (measured, tools/kbench.py).
// Input gate of the host API (lzb_decode_batch with a pinned output buffer): the kernel is launched while the input
// blob is still travelling to the device in chunks on a second stream; after every chunk the copy engine writes the
// device offset reached so far to gate[0].  A warp starts a stream once every 128-byte line the stream touches has
// arrived (whole lines: K1 reads the blob through the non-coherent path, so a line must never be fetched before all of
// it is there).  gate[1] = device offset minus blob offset, gate[2] = device offset of the end of the blob.
// gate[3] (0 = none): device-visible address of one 32-bit "done" word per stream in pinned host memory.  When a stream
// is finished its warp publishes out_len + 1 there (after a system-scope fence by every lane that stored output); the
// host polls these words and lets the COPY ENGINE move each finished stream to the caller's buffer (pinned host memory
// or a peer GPU) while the kernel keeps decoding -- the multi-round form of the host API: the kernel itself then never
// stores across PCIe (measured on the north-star batch: the page-mirroring variant runs 16 % slower than the plain one).
// Returns false if the bytes did not arrive within 5 s (LZB_E_INPUT_TIMEOUT: the host then waits for the upload and
// decodes the stream again; it keeps a stalled copy from hanging the GPU).
// gate[4] != 0: the upload follows the launch's QUEUE instead of the blob (one copy per stream, whole 128-byte lines, in
// the order the warps will ask for them), and the watermark counts queue positions: the stream at position `slot` is
// there once gate[0] > slot.  The static prefix -- the first stream of every warp -- then arrives first and in order,
// instead of wherever its bytes happen to lie in the blob.
__device__ __forceinline__ bool input_arrived(const LzbItem* it, const unsigned long long* gate, unsigned int slot) {
    if (it->kind == LZB_ITEM_PRESET || (it->flags & LZB_ITEM_F_IN_FROM_OUT)) return true;
    unsigned long long need = (it->in_off + it->in_len + gate[1] + 127ull) & ~127ull;
    if (need > gate[2]) need = gate[2];
    if (gate[4]) need = (unsigned long long)slot + 1ull;
    const volatile unsigned long long* wm = gate;
    if (*wm < need) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            __nanosleep(400);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 5000000000ull) return false;
        } while (*wm < need);
    }
    __threadfence_system();  // the stream's bytes are read after the watermark
    return true;
}

// GATE: host-API launches (input gate + done words compiled in).  They always use the SCHED form: in the plain-queue form
// any code around decode_item that depends on the stream index makes ptxas treat the loop's exits as divergent and wrap
// every branch of the bit loop in BSSY / BSYNC pairs (10 -> 78 in the kernel, +20 % instructions; tools/check_sass.py
// counts them) -- the host gives launches without a placement plan a trivial static prefix instead (launch_plan).
// LAT: latency form (lzb_decode_core.h, lat_walk) for launches with few streams per SM -- whole literal table in shared
// memory (warp_smem_bytes covers T_LIT + (0x300 << tab_lclp) entries), no global workspace.
template <bool LIT_GLOBAL, bool MIRROR, int WIDE, bool SCHED, bool GATE, bool LAT = false>
__device__ __forceinline__ void decode_loop(const LzbItem* __restrict__ items, const uint32_t* __restrict__ order,
                                            uint32_t n_items, uint32_t n_static, const uint8_t* __restrict__ in_blob,
                                            uint8_t* out_blob, LzbResult* results, unsigned int* counter, uint32_t tab_lclp,
                                            uint32_t warp_smem_bytes, uint16_t* ws, unsigned long long ws_stride_u16,
                                            const LzbKC& kc, const unsigned long long* gate) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    // broadcast from lane 0 so that the compiler's divergence analysis sees the warp index (and with it every
    // shared-memory table address and every value loaded through it) as warp-uniform: uniform branches then need
    // no BSSY/BSYNC reconvergence pairs in the bit loop
    const int warp = __shfl_sync(FULL_MASK, (int)(threadIdx.x >> 5), 0);
    uint16_t* T = reinterpret_cast<uint16_t*>(smem + (size_t)warp * warp_smem_bytes);
    const TabSm tab = {(uint32_t)__cvta_generic_to_shared(T)};
    // this warp's slice of the global workspace -- indexed with the broadcast (provably uniform) warp index as well
    uint16_t* gws = ws + ((unsigned long long)blockIdx.x * (blockDim.x >> 5) + (unsigned)warp) * ws_stride_u16;
    bool first = SCHED;
    for (;;) {
        unsigned int slot = 0;
        if (SCHED && first) {
            slot = blockIdx.x * (blockDim.x >> 5) + (unsigned)warp;
            first = false;
        } else {
            if (lane == 0) slot = (SCHED ? n_static : 0u) + atomicAdd(counter, 1u);
            slot = __shfl_sync(FULL_MASK, slot, 0);
        }
        if (slot >= n_items) break;
        const uint32_t idx = order[slot];
        if (SCHED && idx == LZB_ORDER_PARK) break;
        // host API: the input blob may still be on its way.  The verdict of the (called) wait is re-broadcast from lane 0:
        // a call's return value counts as divergent for the compiler, and a divergent `continue` here would put
        // reconvergence barriers (BSSY / BSYNC) around every branch of the bit loop below -- that, not the page stores, is
        // what made the round-1 mirror kernels 16 % slower than the plain ones (profiles/r02_e2e_timeline.md).
        if (GATE && gate && !__shfl_sync(FULL_MASK, (int)input_arrived(items + idx, gate, slot), 0)) {
            if (lane == 0) {
                LzbResult r = {};
                r.code = LZB_E_INPUT_TIMEOUT;
                results[idx] = r;
            }
            continue;
        }
        if constexpr (LIT_GLOBAL) {  // whole literal table in the global workspace (reference layout)
            // raw decoder objects (lzb_raw_*): the literal workspace is the decoder's own state record (lzb_types.h)
            uint16_t* lit = gws;
            if (items[idx].flags & LZB_ITEM_F_CARRY)
                lit = reinterpret_cast<uint16_t*>(reinterpret_cast<LzbCarry*>(items[idx].host_out) + 1) + T_LIT;
            const TabPtr plain = {lit}, matched = {lit + 0x100};
            decode_item<true, MIRROR, WIDE, true>(items + idx, in_blob, out_blob, T, lit, tab, plain, matched, kc, tab_lclp, results + idx, lane);
        } else if constexpr (LAT) {  // reference layout in shared memory: matched columns 0x100 entries behind the plain ones
            const TabSm plain = {tab.a + (uint32_t)T_LIT * 2u};
            const TabSm matched = {tab.a + (uint32_t)T_LIT * 2u + 0x200u};
            decode_item<false, MIRROR, WIDE, false, true>(items + idx, in_blob, out_blob, T, nullptr, tab, plain, matched, kc, tab_lclp, results + idx, lane);
        } else {  // plain columns in shared memory, matched columns in the global workspace
            const TabSm plain = {tab.a + (uint32_t)T_LIT * 2u};
            const TabPtr matched = {gws};
            decode_item<false, MIRROR, WIDE>(items + idx, in_blob, out_blob, T, gws, tab, plain, matched, kc, tab_lclp, results + idx, lane);
        }
        __syncwarp();
        if (GATE && gate) {
            const unsigned long long done = gate[3];
            if (done) {
                __threadfence_system();  // this lane's output stores, before the word that releases them to the copy engine
                __syncwarp();
                if (lane == 0)
                    reinterpret_cast<volatile unsigned int*>(done)[idx] = (unsigned int)results[idx].out_len + 1u;
            }
        }
    }
}

#define LZB_KERNEL_ARGS                                                                                              \
    const LzbItem *__restrict__ items, const uint32_t *__restrict__ order, uint32_t n_items, uint32_t n_static,      \
        const uint8_t *__restrict__ in_blob, uint8_t *out_blob, LzbResult *results, unsigned int *counter,           \
        uint32_t tab_lclp, uint32_t warp_smem_bytes, uint16_t *ws, unsigned long long ws_stride_u16,                 \
        const __grid_constant__ LzbKC kc, const unsigned long long *gate
#define LZB_KERNEL_PASS items, order, n_items, n_static, in_blob, out_blob, results, counter, tab_lclp, warp_smem_bytes, ws, ws_stride_u16, kc, gate
#define LZB_DEFINE_K1(NAME, LIT_GLOBAL, MIRROR, WIDE, SCHED, GATE)                                         \
    extern "C" __global__ void __launch_bounds__(LZB_MAX_WARPS * 32, 1) NAME(LZB_KERNEL_ARGS) {          \
        decode_loop<LIT_GLOBAL, MIRROR, WIDE, SCHED, GATE>(LZB_KERNEL_PASS);                              \
    }

// Device-resident batches (lzb_batch_* / lzb_decode_batch_device): no host I/O code at all.
//   WIDE : 0 lean; 1 ("fill") word-wide run fills for batches that expand > 16x; 2 ("copy") 16-byte vector copies of
//          stored chunks for batches that are mostly stored chunks.  Selected by the host from the framing scan.
//   SCHED: launch with a placement plan (above).
LZB_DEFINE_K1(lzb_decode_kernel, false, false, 0, false, false)
LZB_DEFINE_K1(lzb_decode_fill_kernel, false, false, 1, false, false)
LZB_DEFINE_K1(lzb_decode_copy_kernel, false, false, 2, false, false)
LZB_DEFINE_K1(lzb_decode_sched_kernel, false, false, 0, true, false)
LZB_DEFINE_K1(lzb_decode_sched_fill_kernel, false, false, 1, true, false)
LZB_DEFINE_K1(lzb_decode_sched_copy_kernel, false, false, 2, true, false)
// Host API / peer form (input gate, output leaves the device while the kernel runs), always in the SCHED form:
//   "drain"  done word per finished stream, the copy engine moves it (batches of more than two rounds)
//   "mirror" K1 stores finished 4 KiB pages to the destination itself (up to two rounds: all streams end together)
LZB_DEFINE_K1(lzb_decode_drain_kernel, false, false, 0, true, true)
LZB_DEFINE_K1(lzb_decode_drain_fill_kernel, false, false, 1, true, true)
LZB_DEFINE_K1(lzb_decode_drain_copy_kernel, false, false, 2, true, true)
LZB_DEFINE_K1(lzb_decode_mirror_kernel, false, true, 0, true, true)
LZB_DEFINE_K1(lzb_decode_mirror_fill_kernel, false, true, 1, true, true)
LZB_DEFINE_K1(lzb_decode_mirror_copy_kernel, false, true, 2, true, true)
// .lzma streams with lc+lp > 4: literal table in a per-warp global workspace (ws + warp_id * ws_stride_u16).  Also the
// kernel of the raw decoder objects (lzb_raw_*, work items with LZB_ITEM_F_CARRY): there the literal workspace is the
// decoder's state record, and the small tables are loaded from / written back to it (decode_item's CARRY path).
LZB_DEFINE_K1(lzb_decode_biglit_kernel, true, true, 1, false, true)

// Latency form ("LAT"): launches with at most LZB_LAT_WARPS streams per SM, one round.  Static-prefix
// form with the host-I/O code compiled in (gate == nullptr and host_out == 0 switch it off); own launch bounds, so the
// look-ahead registers never spill.
#define LZB_DEFINE_K1_LAT(NAME, MIRROR)                                                                   \
    extern "C" __global__ void __launch_bounds__(LZB_LAT_WARPS * 32, 1) NAME(LZB_KERNEL_ARGS) {          \
        decode_loop<false, MIRROR, 0, true, true, true>(LZB_KERNEL_PASS);                                \
    }
LZB_DEFINE_K1_LAT(lzb_decode_lat_kernel, false)
LZB_DEFINE_K1_LAT(lzb_decode_lat_mirror_kernel, true)

// ------------------------------------------------------------------------------------------------
// K2: per-stream scan -> work items (+ size summary).  One thread per stream.
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void lzb_scan_kernel(int fmt, lzb_options opt, const uint8_t* __restrict__ in_blob,
                                           const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ out_off,
                                           uint32_t n, LzbItem* items, LzbScan* scan, uint64_t mirror_base,
                                           uint64_t* capacity) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = in_blob + in_off[i];
    const uint64_t len = in_off[i + 1] - in_off[i];
    LzbItem it;
    LzbScan sc;
    it.in_off = in_off[i];
    it.in_len = len;
    it.out_off = out_off ? out_off[i] : 0;
    it.out_cap = out_off ? out_off[i + 1] - out_off[i] : 0;
    it.unpacked = LZB_UNKNOWN_SIZE;
    it.memlimit = opt.has_memlimit ? opt.memlimit : ~0ull;
    it.dict_size = 0;
    it.kind = LZB_ITEM_LZMA2;
    it.lc = it.lp = it.pb = 0;
    it.hdr_len = 0;
    it.preset_code = 0;
    it.preset_a0 = 0;
    it.flags = 0;
    it.host_out = 0;
    sc.unpacked = 0;
    sc.flags = 0;
    sc.max_lclp = 0;
    sc.pad[0] = sc.pad[1] = sc.pad[2] = 0;
    sc.stored = 0;

    if (fmt == LZB_FMT_LZMA) {  // LzmaParams::read_header, lzma.rs:96-161
        const uint32_t hdr = opt.unpacked_mode == LZB_UNPACKED_USE_PROVIDED ? 5u : 13u;
        it.kind = LZB_ITEM_LZMA;
        if (len < 1) {
            it.kind = LZB_ITEM_PRESET;
            it.preset_code = LZB_E_HEADER_TOO_SHORT;
        } else if (p[0] >= 225) {
            it.kind = LZB_ITEM_PRESET;
            it.preset_code = LZB_E_LZMA_PROPS;
            it.preset_a0 = p[0];
        } else if (len < hdr) {
            it.kind = LZB_ITEM_PRESET;
            it.preset_code = LZB_E_HEADER_TOO_SHORT;
        } else {
            uint32_t props = p[0];
            it.lc = props % 9;
            props /= 9;
            it.lp = props % 5;
            it.pb = props / 5;
            uint32_t ds = (uint32_t)p[1] | ((uint32_t)p[2] << 8) | ((uint32_t)p[3] << 16) | ((uint32_t)p[4] << 24);
            it.dict_size = ds < 0x1000u ? 0x1000u : ds;
            uint64_t hv = LZB_UNKNOWN_SIZE;
            if (hdr == 13) {
                hv = 0;
                for (int k = 7; k >= 0; k--) hv = (hv << 8) | p[5 + k];
            }
            if (opt.unpacked_mode == LZB_UNPACKED_READ_FROM_HEADER)
                it.unpacked = hv;  // 0xFFFF_FFFF_FFFF_FFFF == marker mandatory == unknown
            else
                it.unpacked = opt.has_provided ? opt.provided : LZB_UNKNOWN_SIZE;
            it.in_off += hdr;
            it.in_len -= hdr;
            it.hdr_len = hdr;
            sc.max_lclp = it.lc + it.lp;
            if (it.unpacked != LZB_UNKNOWN_SIZE) {
                sc.unpacked = it.unpacked;
                sc.flags = 1;
            } else {
                sc.unpacked = len * 8 + 4096;  // heuristic; decode reports LZB_E_CAPACITY if too small
            }
        }
    } else {  // LZMA2 chunk-header walk (well-formed framing assumed; the decode kernel re-walks exactly)
        uint64_t q = 0, total = 0;
        uint32_t lclp = 0, maxl = 0;
        for (;;) {
            if (q >= len) break;
            uint32_t status = p[q++];
            if (status == 0) {
                sc.flags |= 1;
                break;
            }
            if (status == 1 || status == 2) {
                sc.flags |= 2;  // has a stored chunk
                if (len - q < 2) break;
                uint64_t nb = (((uint32_t)p[q] << 8) | p[q + 1]) + 1;
                q += 2;
                if (len - q < nb) break;
                q += nb;
                total += nb;
                sc.stored += nb;
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
            if (lclp > maxl) maxl = lclp;
            total += unpacked;
            if (len - q < packed) break;
            q += packed;
        }
        sc.unpacked = total;
        sc.max_lclp = (uint8_t)maxl;
        if ((sc.flags & 1) && total > 0 && sc.stored == total) {
            it.flags |= LZB_ITEM_F_ALL_STORED;
            it.unpacked = total;
        }
    }
    if (len > 0xFFFFE000ull) {
        it.kind = LZB_ITEM_PRESET;
        it.preset_code = LZB_E_UNSUPPORTED;
    }
    // mirror_base: device-visible address of offset 0 of a second copy of the output blob (pinned host memory or a
    // peer GPU's blob); the mirror variants of K1 store finished pages there (16-byte vectors: both copies aligned)
    if (mirror_base && it.kind != LZB_ITEM_PRESET && (it.out_off & 15) == 0) it.host_out = mirror_base + it.out_off;
    items[i] = it;
    scan[i] = sc;
    if (capacity) {  // lzb_scan's rule for the raw formats (lzb::scan_capacity)
        const uint64_t bound = len * 16384 + (1u << 20);
        uint64_t c;
        if (fmt == LZB_FMT_LZMA)
            c = it.kind == LZB_ITEM_PRESET ? 0 : ((sc.flags & 1) && sc.unpacked <= bound) ? sc.unpacked + 288 : len * 8 + 65536;
        else
            c = sc.unpacked < bound ? sc.unpacked : bound;
        capacity[i] = c;
    }
}

// Output layout on the device: out_off = exclusive prefix sum of the capacities rounded up to 16 bytes (n + 1 entries).
// One CTA; every thread sums a contiguous slice, the slice totals are scanned in shared memory.
extern "C" __global__ void __launch_bounds__(1024) lzb_layout_kernel(const uint64_t* __restrict__ capacity, uint32_t n,
                                                                     uint64_t* out_off) {
    __shared__ uint64_t part[1024];
    const uint32_t t = threadIdx.x, per = (n + 1023u) / 1024u;
    const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
    uint64_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += (capacity[i] + 15ull) & ~15ull;
    part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {  // Hillis-Steele inclusive scan of the slice totals
        const uint64_t v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint64_t run = t ? part[t - 1] : 0;
    for (uint32_t i = lo; i < hi; i++) {
        out_off[i] = run;
        run += (capacity[i] + 15ull) & ~15ull;
    }
    if (t == 1023) out_off[n] = part[1023];
}

// ------------------------------------------------------------------------------------------------
// K3: CRC-32/ISO-HDLC and CRC-64/XZ of byte ranges (crate `crc` catalogue algorithms, src/xz/crc.rs:3-4).
// Phase a: one thread per 4 KiB segment (byte-table CRC from shared memory).
// Phase b: one thread per range folds its segments: crc = shift(crc, 8*seg_len) ^ seg_crc  (GF(2) mulmod).
// ------------------------------------------------------------------------------------------------
#define CRC_SEG 4096u
#define CRC32_POLY 0xEDB88320u
#define CRC64_POLY 0xC96C5795D7870F42ull

struct LzbCrcRange {
    uint64_t off, len;
    uint64_t first_seg;  // index of this range's first segment in the partial arrays
};

extern "C" __global__ void lzb_crc_partial_kernel(const uint8_t* __restrict__ data, const LzbCrcRange* __restrict__ ranges,
                                                  const uint32_t* __restrict__ seg_range, uint64_t n_segs,
                                                  uint32_t* part32, uint64_t* part64) {
    __shared__ uint32_t t32[256];
    __shared__ uint64_t t64[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = i;
        uint64_t e = i;
        for (int k = 0; k < 8; k++) {
            c = (c & 1) ? (c >> 1) ^ CRC32_POLY : c >> 1;
            e = (e & 1) ? (e >> 1) ^ CRC64_POLY : e >> 1;
        }
        t32[i] = c;
        t64[i] = e;
    }
    __syncthreads();
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const LzbCrcRange r = ranges[seg_range[s]];
    const uint64_t k = s - r.first_seg;
    const uint64_t beg = k * CRC_SEG;
    const uint32_t n = (uint32_t)min((uint64_t)CRC_SEG, r.len - beg);
    const uint8_t* p = data + r.off + beg;
    // raw (non-inverted) register from zero: linear in the message, which is what the fold needs
    uint32_t c = 0;
    uint64_t e = 0;
    uint32_t i = 0;
    for (; i < n && ((uintptr_t)(p + i) & 15u); i++) {
        uint8_t b = p[i];
        c = t32[(c ^ b) & 0xFF] ^ (c >> 8);
        e = t64[(e ^ b) & 0xFF] ^ (e >> 8);
    }
    for (; i + 16 <= n; i += 16) {
        uint4 v = *reinterpret_cast<const uint4*>(p + i);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int b8 = 0; b8 < 4; b8++) {
                uint32_t b = (w[j] >> (8 * b8)) & 0xFF;
                c = t32[(c ^ b) & 0xFF] ^ (c >> 8);
                e = t64[(e ^ b) & 0xFF] ^ (e >> 8);
            }
        }
    }
    for (; i < n; i++) {
        uint8_t b = p[i];
        c = t32[(c ^ b) & 0xFF] ^ (c >> 8);
        e = t64[(e ^ b) & 0xFF] ^ (e >> 8);
    }
    part32[s] = c;
    part64[s] = e;
}

// a(x) * b(x) mod P in the reflected representation (bit 31 / bit 63 = x^0)
__device__ __forceinline__ uint32_t mulmod32(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) {
        if (a & 0x80000000u) r ^= b;
        a <<= 1;
        b = (b & 1) ? (b >> 1) ^ CRC32_POLY : b >> 1;
    }
    return r;
}
__device__ __forceinline__ uint64_t mulmod64(uint64_t a, uint64_t b) {
    uint64_t r = 0;
    for (int i = 0; i < 64; i++) {
        if (a & 0x8000000000000000ull) r ^= b;
        a <<= 1;
        b = (b & 1) ? (b >> 1) ^ CRC64_POLY : b >> 1;
    }
    return r;
}
// x^(8*nbytes) mod P
__device__ uint32_t xpow32(uint64_t nbytes) {
    uint32_t r = 0x80000000u, sq = 0x00800000u;  // 1, x^8
    while (nbytes) {
        if (nbytes & 1) r = mulmod32(r, sq);
        sq = mulmod32(sq, sq);
        nbytes >>= 1;
    }
    return r;
}
__device__ uint64_t xpow64(uint64_t nbytes) {
    uint64_t r = 0x8000000000000000ull, sq = 0x0080000000000000ull;
    while (nbytes) {
        if (nbytes & 1) r = mulmod64(r, sq);
        sq = mulmod64(sq, sq);
        nbytes >>= 1;
    }
    return r;
}

extern "C" __global__ void lzb_crc_fold_kernel(const LzbCrcRange* __restrict__ ranges, uint32_t n_ranges,
                                               const uint32_t* __restrict__ part32, const uint64_t* __restrict__ part64,
                                               uint32_t* crc32, uint64_t* crc64) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_ranges) return;
    const LzbCrcRange rg = ranges[r];
    const uint64_t nseg = (rg.len + CRC_SEG - 1) / CRC_SEG;
    const uint32_t f32 = xpow32(CRC_SEG);
    const uint64_t f64 = xpow64(CRC_SEG);
    // register starts at all-ones (init), each segment: reg = reg * x^(8*len) + raw(segment)
    uint32_t c = 0xFFFFFFFFu;
    uint64_t e = ~0ull;
    for (uint64_t k = 0; k < nseg; k++) {
        const uint64_t sl = min((uint64_t)CRC_SEG, rg.len - k * CRC_SEG);
        const uint32_t m32 = sl == CRC_SEG ? f32 : xpow32(sl);
        const uint64_t m64 = sl == CRC_SEG ? f64 : xpow64(sl);
        c = mulmod32(c, m32) ^ part32[rg.first_seg + k];
        e = mulmod64(e, m64) ^ part64[rg.first_seg + k];
    }
    crc32[r] = c ^ 0xFFFFFFFFu;
    crc64[r] = ~e;
}
