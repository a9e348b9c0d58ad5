// lzb_host.cu -- the C ABI (include/lzma_b200.h): contexts, device memory, kernel launches.
// All decoding happens in lzb_kernels.cu on the GPU; the host only plans (lzb_plan.cpp) and moves bytes.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <numeric>
#include <thread>
#include <utility>
#include <chrono>
#include <functional>
#include <vector>

#include "lzb_encode.h"
#include "lzb_plan.h"
#include "lzb_sched.h"
#include "lzb_types.h"

struct LzbCrcRange {
    uint64_t off, len, first_seg;
};
#define CRC_SEG 4096u

#define LZB_K1_PROTO(NAME)                                                                                       \
    extern "C" __global__ void NAME(const LzbItem*, const uint32_t*, uint32_t, uint32_t, const uint8_t*, uint8_t*, \
                                    LzbResult*, unsigned int*, uint32_t, uint32_t, uint16_t*, unsigned long long, \
                                    const LzbKC, const unsigned long long*)
LZB_K1_PROTO(lzb_decode_kernel);
LZB_K1_PROTO(lzb_decode_fill_kernel);
LZB_K1_PROTO(lzb_decode_copy_kernel);
LZB_K1_PROTO(lzb_decode_sched_kernel);
LZB_K1_PROTO(lzb_decode_sched_fill_kernel);
LZB_K1_PROTO(lzb_decode_sched_copy_kernel);
LZB_K1_PROTO(lzb_decode_drain_kernel);
LZB_K1_PROTO(lzb_decode_drain_fill_kernel);
LZB_K1_PROTO(lzb_decode_drain_copy_kernel);
LZB_K1_PROTO(lzb_decode_mirror_kernel);
LZB_K1_PROTO(lzb_decode_mirror_fill_kernel);
LZB_K1_PROTO(lzb_decode_mirror_copy_kernel);
LZB_K1_PROTO(lzb_decode_biglit_kernel);
LZB_K1_PROTO(lzb_decode_lat_kernel);
LZB_K1_PROTO(lzb_decode_lat_mirror_kernel);
extern "C" __global__ void lzb_scan_kernel(int, lzb_options, const uint8_t*, const uint64_t*, const uint64_t*, uint32_t,
                                           LzbItem*, LzbScan*, uint64_t, uint64_t*);
extern "C" __global__ void lzb_layout_kernel(const uint64_t*, uint32_t, uint64_t*);
extern "C" __global__ void lzb_crc_partial_kernel(const uint8_t*, const LzbCrcRange*, const uint32_t*, uint64_t,
                                                  uint32_t*, uint64_t*);
extern "C" __global__ void lzb_stored_decode_kernel(const LzbItem*, const uint32_t*, const uint8_t*, uint8_t*, LzbResult*);
extern "C" __global__ void lzb_store_kernel(const LzbEncItem*, const uint32_t*, const uint32_t*, const uint8_t*, uint8_t*,
                                            uint32_t);
extern "C" __global__ void lzb_frame_kernel(const LzbEncItem*, uint32_t, int, uint8_t*, const LzbXzHead, LzbEncResult*);
extern "C" __global__ void lzb_literal_kernel(const LzbEncItem*, uint32_t, const uint8_t*, uint8_t*,
                                              const lzb_compress_options, LzbEncResult*);
extern "C" __global__ void lzb_crc_fold_kernel(const LzbCrcRange*, uint32_t, const uint32_t*, const uint64_t*,
                                               uint32_t*, uint64_t*);

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max<size_t>(n + n / 8, 4096);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

}  // namespace

struct lzb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // host API input gate (lzb_decode_batch): the input blob is uploaded in chunks on `copy_stream` while K1 already runs
    // on `stream`; after each chunk the offset reached is copied from h_marks to d_gate[0] (see input_arrived())
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t gate_ready = nullptr;
    unsigned long long* h_marks = nullptr;  // pinned: [0..3] initial gate words, [8..] one watermark per chunk
    // drain mode of the host API (multi-round batches): K1 publishes one "done" word per stream into h_done (pinned), the
    // host polls them and the copy engine moves finished streams to the caller's buffer on drain_stream
    cudaStream_t drain_stream = nullptr;
    cudaEvent_t kernels_done = nullptr;
    cudaEvent_t upload_done = nullptr;  // recorded on copy_stream behind the last chunk of a gated upload
    unsigned int* h_done = nullptr;
    size_t h_done_cap = 0;
    // lzb_decode_batch_device keeps ONE batch object alive between calls: its device buffers are reused instead of
    // paying six cudaMalloc / cudaFree (each a device synchronisation) per call
    struct lzb_batch* oneshot = nullptr;
    struct lzb_batch* peer_batch = nullptr;  // lzb_decode_batch_peer's reusable batch (under ctx->mu)
    std::mutex oneshot_mu;
    int sm_count = 0;
    int smem_optin = 0;
    char err[320] = {0};
    std::mutex mu;
    std::vector<std::pair<void*, void*>> ipc_open;  // (pointer handed out, base returned by cudaIpcOpenMemHandle)
    DevBuf d_in, d_out, d_items, d_results, d_order, d_counter, d_scan, d_off, d_crc_ranges, d_crc_segmap, d_crc_part32,
        d_crc_part64, d_crc_out32, d_crc_out64, d_litws, d_matchws, d_gate, d_enc_items, d_enc_results, d_enc_pieces, d_cap;
    int enc_smem_configured = 0;
};
#define LZB_GATE_CHUNK (8ull << 20)  // upload granularity of the gated host path
#define LZB_GATE_MAX_CHUNKS 56
#define LZB_GATE_MARKS 4096          // watermark values of one queue-order upload (h_marks[8 ..])

#define CUDA_TRY(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                     __LINE__);                                                                     \
            return e_ == cudaErrorMemoryAllocation ? LZB_RC_OOM : LZB_RC_CUDA;                       \
        }                                                                                           \
    } while (0)

namespace {

// LZB_TRACE=1: one line per host-API decode on stderr with the host-side timeline (ms since the call started) and the
// in-situ duration of the K1 launch (CUDA events) -- how the end-to-end time of lzb_decode_batch splits up.
struct Trace {
    bool on = getenv("LZB_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double plan = 0, launched = 0, uploaded = 0, synced = 0, kernel_ms = 0;
    double now() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

struct LaunchCfg {
    uint32_t lclp, warp_bytes, warps, grid;
};

// K1 variants: [plain | sched | drain | mirror][lean | fill | copy] (+ the lc+lp > 4 kernel); drain and mirror are the
// host-I/O forms (input gate; done words / page stores) and always take a static prefix like sched.  Their dynamic
// shared-memory limit is raised
// once per device in lzb_create, so launches never touch function attributes (several threads may launch prepared
// batches of one context at the same time).
typedef void (*k1_t)(const LzbItem*, const uint32_t*, uint32_t, uint32_t, const uint8_t*, uint8_t*, LzbResult*,
                     unsigned int*, uint32_t, uint32_t, uint16_t*, unsigned long long, const LzbKC,
                     const unsigned long long*);
const k1_t k1_variants[15] = {
    lzb_decode_kernel,        lzb_decode_fill_kernel,        lzb_decode_copy_kernel,
    lzb_decode_sched_kernel,  lzb_decode_sched_fill_kernel,  lzb_decode_sched_copy_kernel,
    lzb_decode_drain_kernel,  lzb_decode_drain_fill_kernel,  lzb_decode_drain_copy_kernel,
    lzb_decode_mirror_kernel, lzb_decode_mirror_fill_kernel, lzb_decode_mirror_copy_kernel,
    lzb_decode_biglit_kernel,
    lzb_decode_lat_kernel,    lzb_decode_lat_mirror_kernel};  // latency form (few streams per SM): 13 plain, 14 page stores

const char* const k1_names[15] = {
    "lzb_decode_kernel",        "lzb_decode_fill_kernel",        "lzb_decode_copy_kernel",
    "lzb_decode_sched_kernel",  "lzb_decode_sched_fill_kernel",  "lzb_decode_sched_copy_kernel",
    "lzb_decode_drain_kernel",  "lzb_decode_drain_fill_kernel",  "lzb_decode_drain_copy_kernel",
    "lzb_decode_mirror_kernel", "lzb_decode_mirror_fill_kernel", "lzb_decode_mirror_copy_kernel",
    "lzb_decode_biglit_kernel", "lzb_decode_lat_kernel",         "lzb_decode_lat_mirror_kernel"};

LaunchCfg decode_config(const lzb_ctx* ctx, uint32_t n, uint32_t lclp) {
    LaunchCfg c;
    c.lclp = lclp;
    c.warp_bytes = (lzb_table_u16(lclp) * 2 + 15u) & ~15u;
    uint32_t max_warps = std::min<uint32_t>(LZB_MAX_WARPS, (uint32_t)ctx->smem_optin / c.warp_bytes);
    if (max_warps < 1) max_warps = 1;
    // Streams are indivisible and (in a batch of similar streams) take about the same time, so the kernel runs in
    // "rounds" of sm_count * warps streams.  Use the fewest rounds the residency limit allows and then the smallest
    // warp count that still fits them: e.g. 4 096 streams on 148 SMs -> 1 round of 28 warps instead of 1.15 rounds
    // of 24 (whose second round would leave most of the chip idle).
    const uint32_t sm = (uint32_t)ctx->sm_count;
    const uint32_t rounds = std::max<uint32_t>(1, (n + sm * max_warps - 1) / (sm * max_warps));
    c.warps = std::max<uint32_t>(1, std::min<uint32_t>(max_warps, (n + sm * rounds - 1) / (sm * rounds)));
    c.grid = std::max<uint32_t>(1, std::min<uint32_t>(sm, (n + c.warps - 1) / c.warps));
    return c;
}

// One decode = up to two launches: streams whose tables fit shared memory (every LZMA2 stream; .lzma with
// lc+lp <= 4) and .lzma streams with lc+lp > 4 whose literal table goes to a global workspace.
struct DecodePlan {
    std::vector<uint32_t> order_small, order_big;  // item indices, longest compressed stream first
    std::vector<uint32_t> order_stored;            // stored-chunk-only LZMA2 streams: copy kernel, no range decoder
    LaunchCfg cfg_small{}, cfg_big{};
    uint32_t n_small = 0;         // streams of the first launch (order_small may also hold LZB_ORDER_PARK entries)
    uint32_t n_static = 0;        // leading order_small entries that are pre-assigned first items (lzb_sched.h)
    uint32_t parked = 0;          // warps the placement keeps out of the launch
    uint64_t big_stride_u16 = 0;  // workspace u16 per warp
    int wide = 0;                 // K1 variant: 0 lean, 1 word-wide run fills, 2 vector stored-chunk copies
    bool lat = false;             // first launch with the latency kernels (at most LZB_LAT_WARPS streams per SM, one round)
    bool host_io = false;         // launch with the host-I/O kernels (input gate; done words or page stores)
};

void make_plan(const lzb_ctx* ctx, const LzbItem* items, uint32_t n, uint32_t lzma2_lclp_hint, uint64_t stored_bytes,
               DecodePlan* p, bool route_stored = true, bool host_io = false) {
    p->host_io = host_io;
    uint32_t lclp_small = 0, lclp_big = 0;
    // stored-chunk-only streams whose output fits go to the copy kernel (not with a host mirror: that kernel does not
    // stream pages to the host); everything below plans K1 for the rest
    auto stored_route = [&](const LzbItem& it) {
        return route_stored && it.kind == LZB_ITEM_LZMA2 && (it.flags & LZB_ITEM_F_ALL_STORED) &&
               !(it.flags & LZB_ITEM_F_IN_FROM_OUT) && it.unpacked <= it.out_cap;
    };
    p->order_stored.clear();
    // K1 variants beside the lean default (which is 3-6 % faster on everything else: instruction-cache footprint):
    // vector stored-chunk copies when stored chunks carry >= 80 % of the batch's output (break-even measured at ~5x the
    // range-coded bytes: C3 with 3 stored chunks in 8 192 streams lost 6 % to the copy code it never needed), word-wide
    // run fills when the batch expands so much (> 16x) that it must be long runs
    uint64_t in_sum = 0, cap_sum = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (stored_route(items[i])) {
            stored_bytes -= std::min<uint64_t>(stored_bytes, items[i].unpacked);
            continue;
        }
        in_sum += items[i].in_len;
        cap_sum += items[i].out_cap;
    }
    p->wide = stored_bytes * 5 >= cap_sum * 4 && stored_bytes ? 2 : cap_sum > 16 * in_sum ? 1 : 0;
    p->order_small.clear();
    p->order_big.clear();
    const bool force_big = getenv("LZB_FORCE_BIGLIT") != nullptr;  // test switch: every .lzma stream through the lc+lp > 4 kernel
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t lclp = (uint32_t)items[i].lc + items[i].lp;
        if (stored_route(items[i])) {
            p->order_stored.push_back(i);
        } else if (items[i].kind == LZB_ITEM_LZMA && (lclp > 4 || force_big)) {
            p->order_big.push_back(i);
            lclp_big = std::max(lclp_big, lclp);
        } else {
            p->order_small.push_back(i);
            if (items[i].kind == LZB_ITEM_LZMA) lclp_small = std::max(lclp_small, lclp);
            if (items[i].kind == LZB_ITEM_LZMA2) lclp_small = std::max(lclp_small, std::min<uint32_t>(lzma2_lclp_hint, 4));
        }
    }
    auto by_len = [&](uint32_t a, uint32_t b) { return items[a].in_len > items[b].in_len; };
    std::stable_sort(p->order_small.begin(), p->order_small.end(), by_len);
    std::stable_sort(p->order_big.begin(), p->order_big.end(), by_len);
    p->n_small = (uint32_t)p->order_small.size();
    p->cfg_small = decode_config(ctx, p->n_small, lclp_small);
    p->n_static = p->parked = 0;
    p->lat = false;
    if (p->n_small && p->wide == 0 && !getenv("LZB_NO_LAT")) {
        // Few streams per SM: a stream's time is the dependent chain of its decisions, not the instruction count.  The
        // latency kernels (lzb_decode_core.h, lat_walk) fetch probabilities ahead of the decisions that select them and keep
        // the whole literal table in shared memory; they win while a warp has most of an SM sub-partition to itself.
        const uint32_t lat_bytes = (lzb_lat_table_u16(lclp_small) * 2 + 15u) & ~15u;
        uint32_t lat_warps = std::min<uint32_t>(LZB_LAT_WARPS, (uint32_t)ctx->smem_optin / lat_bytes);
        if (const char* e = getenv("LZB_LAT_WARPS_MAX")) lat_warps = std::min<uint32_t>(lat_warps, (uint32_t)atoi(e));  // experiments
        const uint32_t sm = (uint32_t)ctx->sm_count;
        if (lat_warps && p->n_small <= sm * lat_warps) {
            p->lat = true;
            LaunchCfg& c = p->cfg_small;
            c.warp_bytes = lat_bytes;
            c.warps = (p->n_small + sm - 1) / sm;
            c.grid = std::min<uint32_t>(sm, (p->n_small + c.warps - 1) / c.warps);
        }
    }
    if (p->n_small > (uint32_t)ctx->sm_count * p->cfg_small.warps) {
        // several rounds: streams that would still be running when the queue is empty get less crowded SMs
        std::vector<double> work(n);
        for (uint32_t i = 0; i < n; i++) work[i] = (double)items[i].in_len;
        lzb_sched::Plan sp = lzb_sched::plan(p->order_small, work, (uint32_t)ctx->sm_count, p->cfg_small.warps);
        if (sp.throttled) {
            p->order_small.swap(sp.order);
            p->n_static = sp.n_static;
            p->parked = sp.parked;
            p->cfg_small.grid = sp.grid;
        }
    }
    if ((host_io || p->lat) && !p->n_static && p->n_small) {
        // the host-I/O kernels exist in the static-prefix form only (lzb_kernels.cu): without a placement plan, hand the
        // head of the queue out round-robin over the CTAs, which is what the dynamic queue does in effect
        const uint32_t grid = p->cfg_small.grid, warps = p->cfg_small.warps, ns = grid * warps;
        std::vector<uint32_t> order(ns, LZB_ORDER_PARK);
        for (uint32_t c = 0; c < grid; c++)
            for (uint32_t w = 0; w < warps; w++)
                if ((uint64_t)w * grid + c < p->n_small) order[c * warps + w] = p->order_small[w * grid + c];
        if (p->n_small > ns) order.insert(order.end(), p->order_small.begin() + ns, p->order_small.end());
        p->order_small.swap(order);
        p->n_static = ns;
    }
    if (!p->order_big.empty()) {
        LaunchCfg& c = p->cfg_big;
        c.lclp = lclp_big;
        c.warp_bytes = ((uint32_t)T_LIT * 2 + 15u) & ~15u;
        p->big_stride_u16 = (uint64_t)0x300u << lclp_big;
        const uint64_t ws_budget = 1ull << 30;  // bound the workspace to 1 GiB
        uint32_t total_warps = (uint32_t)std::min<uint64_t>(p->order_big.size(), (uint64_t)ctx->sm_count * LZB_MAX_WARPS);
        total_warps = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(total_warps, ws_budget / (p->big_stride_u16 * 2)));
        c.warps = std::max<uint32_t>(1, std::min<uint32_t>(LZB_MAX_WARPS, (total_warps + ctx->sm_count - 1) / ctx->sm_count));
        c.grid = std::max<uint32_t>(1, total_warps / c.warps);
    }
}

// Index into k1_variants / k1_names of the first launch's kernel.
int k1_variant(const DecodePlan& p, bool mirror) {
    return p.lat ? (mirror ? 14 : 13) : (p.host_io ? (mirror ? 9 : 6) : p.n_static ? 3 : 0) + p.wide;
}

// Enqueues counter resets + K1 launch(es) on `s`.  All pointers are device pointers; d_order holds order_small
// followed by order_big; d_counter holds two counters.
int launch_plan(lzb_ctx* ctx, cudaStream_t s, const DecodePlan& p, const LzbItem* d_items, const uint32_t* d_order,
                const uint8_t* d_in_base, uint8_t* d_out_base, LzbResult* d_results, unsigned int* d_counter,
                bool mirror = false, const unsigned long long* d_gate = nullptr,
                std::function<int(int)>* upload = nullptr, int* upload_phase = nullptr, DevBuf* own_matchws = nullptr,
                DevBuf* own_litws = nullptr) {
    // per-warp workspaces: the ctx's own for the host API (calls are serialised by ctx->mu), the batch's own for prepared
    // device batches, which may be in flight on different streams at the same time
    DevBuf& matchws = own_matchws ? *own_matchws : ctx->d_matchws;
    DevBuf& litws = own_litws ? *own_litws : ctx->d_litws;
    CUDA_TRY(ctx, cudaMemsetAsync(d_counter, 0, 2 * sizeof(unsigned int), s));
    // everything small this launch needs has been enqueued: now the gated upload of the input blob may occupy the copy
    // engine.  It is enqueued BEFORE the kernel so that a launch that blocks the host (profilers, compute-sanitizer,
    // CUDA_LAUNCH_BLOCKING) cannot wait for bytes nobody has sent yet.
    auto fire = [&]() -> int {  // phase 0 of the gated upload; the caller runs phase 1 behind the launches (upload_rest)
        if (!upload || !*upload || !upload_phase || *upload_phase != 0) return LZB_RC_OK;
        *upload_phase = 1;
        return (*upload)(0);
    };
    const uint32_t ns = (uint32_t)p.order_small.size(), nb = (uint32_t)p.order_big.size();
    const uint32_t nst = (uint32_t)p.order_stored.size();
    const LzbKC kc = LZB_KC_INIT;
    if (ns) {
        const LaunchCfg& c = p.cfg_small;
        const int smem = (int)(c.warps * c.warp_bytes);
        // per-warp global workspace for the matched-literal columns (L2-resident: 8 KiB per warp at lc+lp = 3)
        const uint64_t mstride = p.lat ? 0 : lzb_matched_u16(c.lclp);
        if (!p.lat) CUDA_TRY(ctx, matchws.ensure((size_t)c.grid * c.warps * mstride * 2));
        const k1_t* kernels = k1_variants;
        const int v = k1_variant(p, mirror);
        if (int rc = fire()) return rc;
        kernels[v]<<<c.grid, c.warps * 32, smem, s>>>(d_items, d_order, ns, p.n_static, d_in_base, d_out_base, d_results,
                                                      d_counter, c.lclp, c.warp_bytes, matchws.as<uint16_t>(), mstride, kc,
                                                      d_gate);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (nb) {
        const LaunchCfg& c = p.cfg_big;
        const int smem = (int)(c.warps * c.warp_bytes);
        CUDA_TRY(ctx, litws.ensure((size_t)c.grid * c.warps * p.big_stride_u16 * 2));
        if (int rc = fire()) return rc;
        lzb_decode_biglit_kernel<<<c.grid, c.warps * 32, smem, s>>>(d_items, d_order + ns, nb, 0u, d_in_base, d_out_base,
                                                                    d_results, d_counter + 1, c.lclp, c.warp_bytes,
                                                                    litws.as<uint16_t>(), p.big_stride_u16, kc, d_gate);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (nst) {
        if (int rc = fire()) return rc;
        // the stored-chunk copy kernel has no input gate: it starts behind the last chunk of an upload still in flight
        if (d_gate) CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->upload_done, 0));
        lzb_stored_decode_kernel<<<nst, 256, 0, s>>>(d_items, d_order + ns + nb, d_in_base, d_out_base, d_results);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return LZB_RC_OK;
}

// Phase 1 of a gated upload whose phase 0 ran inside launch_plan.
int upload_rest(std::function<int(int)>* upload, int* upload_phase) {
    if (!upload || !*upload || !upload_phase || *upload_phase != 1) return LZB_RC_OK;
    *upload_phase = 2;
    return (*upload)(1);
}

int upload_order(lzb_ctx* ctx, cudaStream_t s, const DecodePlan& p, DevBuf& d_order) {
    const size_t ns = p.order_small.size(), nb = p.order_big.size(), nst = p.order_stored.size();
    CUDA_TRY(ctx, d_order.ensure((ns + nb + nst) * 4 + 4));
    if (nst)
        CUDA_TRY(ctx, cudaMemcpyAsync(d_order.as<uint32_t>() + ns + nb, p.order_stored.data(), nst * 4, cudaMemcpyHostToDevice, s));
    if (ns) CUDA_TRY(ctx, cudaMemcpyAsync(d_order.p, p.order_small.data(), ns * 4, cudaMemcpyHostToDevice, s));
    if (nb) CUDA_TRY(ctx, cudaMemcpyAsync(d_order.as<uint32_t>() + ns, p.order_big.data(), nb * 4, cudaMemcpyHostToDevice, s));
    return LZB_RC_OK;
}

// Arms K1's input gate for a blob of `in_bytes` bytes at `src` (pinned host memory, or a peer GPU's memory) that goes to
// ctx->d_in + lead: the gate words are uploaded on ctx->stream, and `*upload` becomes the closure that enqueues the
// chunked copy (+ one watermark update per chunk) on ctx->copy_stream.  The closure is run by launch_plan right before
// the first K1 launch is enqueued.
int arm_gate(lzb_ctx* ctx, const uint8_t* src, uint64_t lead, uint64_t in_lo, uint64_t in_bytes,
             std::function<int(int)>* upload, const unsigned long long** d_gate) {
    uint64_t chunk = LZB_GATE_CHUNK;
    while ((lead + in_bytes + chunk - 1) / chunk > LZB_GATE_MAX_CHUNKS) chunk *= 2;
    CUDA_TRY(ctx, ctx->d_gate.ensure(64));
    unsigned long long* hm = ctx->h_marks;
    hm[0] = 0;                                      // watermark: device offset reached so far
    hm[1] = (unsigned long long)(lead - in_lo);     // device offset = blob offset + this (mod 2^64)
    hm[2] = (unsigned long long)(lead + in_bytes);  // device offset of the end of the blob
    hm[3] = 0;                                      // no done words (set_done_words)
    hm[4] = 0;                                      // the watermark counts bytes (arm_gate_queue: queue positions)
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_gate.p, hm, 40, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->gate_ready, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->gate_ready, 0));
    *upload = [=](int phase) -> int {  // phase 0: before the first kernel launch (everything); phase 1: behind it (nothing)
        if (phase) return LZB_RC_OK;
        uint32_t k = 0;
        for (uint64_t lo = lead; lo < lead + in_bytes; k++) {  // chunk boundaries at device offsets k * chunk
            const uint64_t hi = std::min<uint64_t>((lo / chunk + 1) * chunk, lead + in_bytes);
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_in.as<uint8_t>() + lo, src + (lo - lead), hi - lo, cudaMemcpyDefault,
                                          ctx->copy_stream));
            hm[8 + k] = hi;
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_gate.p, &hm[8 + k], 8, cudaMemcpyHostToDevice, ctx->copy_stream));
            lo = hi;
        }
        CUDA_TRY(ctx, cudaEventRecord(ctx->upload_done, ctx->copy_stream));
        return LZB_RC_OK;
    };
    *d_gate = ctx->d_gate.as<unsigned long long>();
    return LZB_RC_OK;
}

// Queue-order form of the gated upload (raw formats, one work item per stream): re-arms the gate armed by arm_gate so that
// the watermark counts positions of `queue` (the launch's order array; LZB_ORDER_PARK entries count), and replaces the
// upload closure.  Every stream is one copy of whole 128-byte lines of the device blob (K1 reads its input through the
// non-coherent path: a line must be complete before anything on it is read; neighbours share their boundary lines, which
// are simply copied twice with the same bytes).  Phase 0 (before the kernel launch) enqueues `tail` (streams that are not
// in K1's queue: the stored-chunk copy kernel's, which has no gate and waits for ctx->upload_done -- that event has to be
// recorded BEFORE the launch is enqueued, so these go first) and the head of the queue; phase 1 (behind the launch, so
// that the launch is not held up by ~10^5 driver calls) the rest.  Returns 1 without touching the gate when the tail is
// too large to go first (the caller keeps the blob-order upload).
int arm_gate_queue(lzb_ctx* ctx, const uint8_t* src, uint64_t lead, uint64_t in_lo, uint64_t in_bytes,
                   const LzbItem* items, const std::vector<uint32_t>& queue, const std::vector<uint32_t>& tail,
                   std::function<int(int)>* upload) {
    uint64_t tail_bytes = 0;
    for (uint32_t i : tail) tail_bytes += items[i].in_len;
    if (tail_bytes > (64ull << 20)) return 1;
    unsigned long long* hm = ctx->h_marks;
    hm[6] = 1;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_gate.as<unsigned long long>() + 4, &hm[6], 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->gate_ready, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->gate_ready, 0));
    const size_t npos = queue.size();
    // Environments in which a kernel launch blocks the host until the kernel has finished (profilers and
    // compute-sanitizer inject themselves through CUDA_INJECTION64_PATH; CUDA_LAUNCH_BLOCKING): nothing can be enqueued
    // behind the launch, so the whole upload goes in front of it.
    const char* lb = getenv("CUDA_LAUNCH_BLOCKING");
    const bool blocking_launch = getenv("CUDA_INJECTION64_PATH") || getenv("NV_NSIGHT_INJECTION_PORT_BASE") ||
                                 getenv("NV_TPS_LAUNCH_TOKEN") || (lb && lb[0] && lb[0] != '0');
    const size_t head = blocking_launch ? npos : std::min<size_t>(npos, 768);
    // a watermark update every `step` positions: at most LZB_GATE_MARKS of them
    const size_t step = std::max<size_t>(16, (npos + LZB_GATE_MARKS - 2) / (LZB_GATE_MARKS - 1));
    std::vector<uint32_t> q(queue), t(tail);
    *upload = [=](int phase) -> int {
        auto copy_stream_of = [&](uint32_t i) -> int {
            const LzbItem& it = items[i];
            if (it.kind == LZB_ITEM_PRESET || (it.flags & LZB_ITEM_F_IN_FROM_OUT) || !it.in_len) return LZB_RC_OK;
            // device offsets of the stream (items carry blob offsets; hdr_len bytes in front belong to it too)
            uint64_t lo = lead + (it.in_off - it.hdr_len - in_lo), hi = lead + (it.in_off + it.in_len - in_lo);
            lo = std::max<uint64_t>(lead, lo & ~127ull);
            hi = std::min<uint64_t>(lead + in_bytes, (hi + 127ull) & ~127ull);
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_in.as<uint8_t>() + lo, src + (lo - lead), hi - lo, cudaMemcpyDefault,
                                          ctx->copy_stream));
            return LZB_RC_OK;
        };
        if (!phase) {
            for (uint32_t i : t)
                if (int rc = copy_stream_of(i)) return rc;
            CUDA_TRY(ctx, cudaEventRecord(ctx->upload_done, ctx->copy_stream));
        }
        const size_t from = phase ? head : 0, to = phase ? npos : head;
        for (size_t k = from; k < to; k++) {
            if (q[k] != LZB_ORDER_PARK)
                if (int rc = copy_stream_of(q[k])) return rc;
            if ((k + 1) % step == 0 || k + 1 == npos) {
                unsigned long long* mark = &hm[8 + (k + step) / step];  // one pinned word per update, never reused
                *mark = k + 1;
                CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_gate.p, mark, 8, cudaMemcpyHostToDevice, ctx->copy_stream));
            }
        }
        return LZB_RC_OK;
    };
    return LZB_RC_OK;
}

// Drain mode: n zeroed done words in pinned host memory, their address in gate[3].  *d_gate may still be null (input
// already on the device): then a gate that is open from the start is set up.
int set_done_words(lzb_ctx* ctx, uint32_t n, const unsigned long long** d_gate) {
    if (n > ctx->h_done_cap) {
        if (ctx->h_done) cudaFreeHost(ctx->h_done);
        ctx->h_done = nullptr;
        ctx->h_done_cap = 0;
        const size_t want = (size_t)n + n / 4 + 1024;
        CUDA_TRY(ctx, cudaHostAlloc((void**)&ctx->h_done, want * sizeof(unsigned int), cudaHostAllocDefault));
        ctx->h_done_cap = want;
    }
    memset(ctx->h_done, 0, (size_t)n * sizeof(unsigned int));
    unsigned long long* hm = ctx->h_marks;
    CUDA_TRY(ctx, ctx->d_gate.ensure(64));
    if (!*d_gate) {
        hm[0] = ~0ull;  // everything has arrived
        hm[1] = hm[2] = hm[3] = hm[4] = 0;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_gate.p, hm, 40, cudaMemcpyHostToDevice, ctx->stream));
        *d_gate = ctx->d_gate.as<unsigned long long>();
    }
    hm[5] = (unsigned long long)(uintptr_t)ctx->h_done;  // UVA: pinned host memory has the same address on the device
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_gate.as<unsigned long long>() + 3, &hm[5], 8, cudaMemcpyHostToDevice, ctx->stream));
    return LZB_RC_OK;
}

// Drain mode, host side: while the kernels recorded in ctx->kernels_done run, poll the done words in schedule order and
// let the copy engine move every finished stream from the device blob to `dst_base` (same offsets; pinned host memory
// or a peer GPU).  `order` is the launch's queue (may hold LZB_ORDER_PARK); copied[i] is set for every stream moved.
// Streams complete roughly in queue order, so a sliding window over the queue is enough.  Returns after the kernels
// have finished and every published stream has been enqueued; the caller handles the rest (kernels without done words)
// and synchronises ctx->drain_stream.
int drain_finished(lzb_ctx* ctx, const std::vector<uint32_t>& order, const LzbItem* items, const uint8_t* d_out_base,
                   uint8_t* dst_base, std::vector<uint8_t>& copied, uint32_t window) {
    volatile unsigned int* done = ctx->h_done;
    size_t frontier = 0;
    auto sweep = [&](size_t hi) -> int {
        int moved = 0;
        for (size_t k = frontier; k < hi; k++) {
            const uint32_t i = order[k];
            if (i == LZB_ORDER_PARK || copied[i]) continue;
            const unsigned int f = done[i];
            if (!f) continue;
            copied[i] = 1;
            moved++;
            if (f > 1 && !(items[i].flags & LZB_ITEM_F_OUT_SCRATCH))
                CUDA_TRY(ctx, cudaMemcpyAsync(dst_base + items[i].out_off, d_out_base + items[i].out_off, f - 1,
                                              cudaMemcpyDefault, ctx->drain_stream));
        }
        while (frontier < order.size() && (order[frontier] == LZB_ORDER_PARK || copied[order[frontier]])) frontier++;
        return moved;
    };
    for (;;) {
        const bool finished = cudaEventQuery(ctx->kernels_done) == cudaSuccess;
        const int moved = sweep(finished ? order.size() : std::min(order.size(), frontier + window));
        if (moved < 0) return moved;
        if (finished) break;
        if (!moved) std::this_thread::sleep_for(std::chrono::microseconds(40));
    }
    return LZB_RC_OK;
}

// The one shipped Executor: work items run on the GPU.
class CudaExecutor : public lzb::Executor {
   public:
    // host_mirror: device-visible address of offset 0 of the caller's pinned host output (nullptr = none)
    // d_gate: input gate words of a chunked upload still in flight (nullptr = the input is already on the device)
    CudaExecutor(lzb_ctx* ctx, cudaStream_t s, const uint8_t* d_in_base, uint8_t* d_out_base, uint8_t* host_mirror = nullptr,
                 const unsigned long long* d_gate = nullptr)
        : ctx_(ctx), s_(s), in_(d_in_base), out_(d_out_base), hmirror_(host_mirror), gate_(d_gate) {}
    // The gated upload of the input blob.  Phase 0 runs right before the first K1 launch is enqueued (after the small
    // uploads of the plan, which must not queue behind it on the copy engine), phase 1 behind the launch.
    std::function<int(int)> before_first_kernel;
    int upload_phase = 0;
    cudaStream_t copy_stream = nullptr;  // where that upload runs
    // set for the raw formats (one work item per stream): the upload may follow the launch's queue instead of the blob
    struct QueueUpload {
        const uint8_t* src = nullptr;  // first byte of the blob in the caller's memory
        uint64_t lead = 0, in_lo = 0, in_bytes = 0;
    } queue_upload;
    Trace* trace = nullptr;
    bool all_mirrored() const { return hmirror_ != nullptr && !unmirrored_; }

    int decode(const LzbItem* items, uint32_t n, uint32_t max_lclp, uint64_t stored_bytes, LzbResult* results) override {
        int rc = run(items, n, max_lclp, stored_bytes, results);
        if (rc != LZB_RC_OK) return rc;
        rc = redo_timeouts(items, n, max_lclp, stored_bytes, results);
        if (rc != LZB_RC_OK) return rc;
        // the framing scan can under-estimate lc+lp on malformed LZMA2 streams: rerun just those with the maximum
        std::vector<uint32_t> redo;
        for (uint32_t i = 0; i < n; i++)
            if (items[i].kind == LZB_ITEM_LZMA2 && results[i].code == LZB_E_UNSUPPORTED && results[i].a0 <= 4 &&
                results[i].a1 < 4)
                redo.push_back(i);
        if (!redo.empty()) {
            std::vector<LzbItem> sub(redo.size());
            std::vector<LzbResult> subres(redo.size());
            for (size_t k = 0; k < redo.size(); k++) sub[k] = items[redo[k]];
            rc = run(sub.data(), (uint32_t)sub.size(), 4, stored_bytes, subres.data());
            if (rc != LZB_RC_OK) return rc;
            for (size_t k = 0; k < redo.size(); k++) results[redo[k]] = subres[k];
        }
        return LZB_RC_OK;
    }

    int crc(const lzb::CrcRange* ranges, uint32_t n, uint32_t* crc32, uint64_t* crc64) override {
        std::vector<LzbCrcRange> rg(n);
        std::vector<uint32_t> segmap;
        uint64_t nseg = 0;
        for (uint32_t i = 0; i < n; i++) {
            rg[i].off = ranges[i].off;
            rg[i].len = ranges[i].len;
            rg[i].first_seg = nseg;
            uint64_t k = (ranges[i].len + CRC_SEG - 1) / CRC_SEG;
            segmap.insert(segmap.end(), (size_t)k, i);
            nseg += k;
        }
        lzb_ctx* ctx = ctx_;
        CUDA_TRY(ctx, ctx->d_crc_ranges.ensure(n * sizeof(LzbCrcRange)));
        CUDA_TRY(ctx, ctx->d_crc_segmap.ensure(std::max<uint64_t>(nseg, 1) * 4));
        CUDA_TRY(ctx, ctx->d_crc_part32.ensure(std::max<uint64_t>(nseg, 1) * 4));
        CUDA_TRY(ctx, ctx->d_crc_part64.ensure(std::max<uint64_t>(nseg, 1) * 8));
        CUDA_TRY(ctx, ctx->d_crc_out32.ensure(n * 4));
        CUDA_TRY(ctx, ctx->d_crc_out64.ensure(n * 8));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_crc_ranges.p, rg.data(), n * sizeof(LzbCrcRange), cudaMemcpyHostToDevice, s_));
        if (nseg) {
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_crc_segmap.p, segmap.data(), nseg * 4, cudaMemcpyHostToDevice, s_));
            lzb_crc_partial_kernel<<<(unsigned)((nseg + 127) / 128), 128, 0, s_>>>(
                out_, ctx->d_crc_ranges.as<LzbCrcRange>(), ctx->d_crc_segmap.as<uint32_t>(), nseg,
                ctx->d_crc_part32.as<uint32_t>(), ctx->d_crc_part64.as<uint64_t>());
            CUDA_TRY(ctx, cudaGetLastError());
        }
        lzb_crc_fold_kernel<<<(n + 63) / 64, 64, 0, s_>>>(ctx->d_crc_ranges.as<LzbCrcRange>(), n,
                                                          ctx->d_crc_part32.as<uint32_t>(), ctx->d_crc_part64.as<uint64_t>(),
                                                          ctx->d_crc_out32.as<uint32_t>(), ctx->d_crc_out64.as<uint64_t>());
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaMemcpyAsync(crc32, ctx->d_crc_out32.p, n * 4, cudaMemcpyDeviceToHost, s_));
        CUDA_TRY(ctx, cudaMemcpyAsync(crc64, ctx->d_crc_out64.p, n * 8, cudaMemcpyDeviceToHost, s_));
        CUDA_TRY(ctx, cudaStreamSynchronize(s_));
        return LZB_RC_OK;
    }

    int scratch(uint64_t bytes, uint64_t* off) override {
        lzb_ctx* ctx = ctx_;
        void* p = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&p, bytes + 64));
        scratch_.push_back(p);
        *off = (uint64_t)((uint8_t*)p - out_);  // output-blob coordinates (pointer difference modulo 2^64)
        return LZB_RC_OK;
    }
    int read_out(uint64_t off, uint64_t len, uint8_t* dst) override {
        lzb_ctx* ctx = ctx_;
        CUDA_TRY(ctx, cudaMemcpyAsync(dst, out_ + off, len, cudaMemcpyDeviceToHost, s_));
        CUDA_TRY(ctx, cudaStreamSynchronize(s_));
        return LZB_RC_OK;
    }
    ~CudaExecutor() override {
        for (void* p : scratch_) cudaFree(p);
    }

   private:
    // Streams whose input had not arrived within the gate's timeout (an environment that stalls the upload while K1
    // runs): wait for the upload, then decode just those -- the gate is open by then.
    int redo_timeouts(const LzbItem* items, uint32_t n, uint32_t lclp_hint, uint64_t stored_bytes, LzbResult* results) {
        std::vector<uint32_t> redo;
        for (uint32_t i = 0; i < n; i++)
            if (results[i].code == LZB_E_INPUT_TIMEOUT) redo.push_back(i);
        if (redo.empty()) return LZB_RC_OK;
        if (copy_stream) CUDA_TRY(ctx_, cudaStreamSynchronize(copy_stream));
        std::vector<LzbItem> sub(redo.size());
        std::vector<LzbResult> subres(redo.size());
        for (size_t k = 0; k < redo.size(); k++) sub[k] = items[redo[k]];
        int rc = run(sub.data(), (uint32_t)sub.size(), lclp_hint, stored_bytes, subres.data());
        if (rc != LZB_RC_OK) return rc;
        for (size_t k = 0; k < redo.size(); k++) results[redo[k]] = subres[k];
        return LZB_RC_OK;
    }
    int run(const LzbItem* items, uint32_t n, uint32_t lclp_hint, uint64_t stored_bytes, LzbResult* results) {
        lzb_ctx* ctx = ctx_;
        DecodePlan plan;
        // Output to a pinned host buffer, two forms.  One round of streams (every stream resident from the start): K1's
        // mirror variants store finished 4 KiB pages to the host buffer themselves (all streams end together, nothing else
        // could overlap the transfer).  More than one round: drain mode -- the plain kernels publish a done word per
        // stream and the copy engine moves finished streams while later ones decode (the kernel never stores across PCIe;
        // DESIGN.md section 5).  LZB_DRAIN_ROUNDS overrides the threshold (in rounds; experiments).
        const char* dr = getenv("LZB_DRAIN_ROUNDS");
        const double drain_rounds = dr ? atof(dr) : 1.0;
        const bool drain = hmirror_ != nullptr && !getenv("LZB_NO_DRAIN") &&
                           (double)n > drain_rounds * (double)ctx->sm_count * LZB_MAX_WARPS;
        const bool mirror = hmirror_ != nullptr && !drain;
        make_plan(ctx, items, n, lclp_hint, stored_bytes, &plan, /*route_stored=*/!mirror, /*host_io=*/hmirror_ != nullptr);
        CUDA_TRY(ctx, ctx->d_items.ensure(n * sizeof(LzbItem)));
        CUDA_TRY(ctx, ctx->d_results.ensure(n * sizeof(LzbResult)));
        CUDA_TRY(ctx, ctx->d_counter.ensure(64));
        const LzbItem* up = items;
        std::vector<LzbItem> patched;
        if (mirror) {  // the mirror copies 16-byte vectors: both copies of a stream's region must be 16-byte aligned
            patched.assign(items, items + n);
            for (auto& it : patched) {
                if (it.flags & LZB_ITEM_F_OUT_SCRATCH) continue;  // intermediate result of a filter chain
                if ((it.out_off & 15) == 0 && it.kind != LZB_ITEM_PRESET)
                    it.host_out = (uint64_t)(uintptr_t)(hmirror_ + it.out_off);
                else if (it.kind != LZB_ITEM_PRESET)
                    unmirrored_ = true;
            }
            up = patched.data();
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_items.p, up, n * sizeof(LzbItem), cudaMemcpyHostToDevice, s_));
        int rc = upload_order(ctx, s_, plan, ctx->d_order);
        if (rc != LZB_RC_OK) return rc;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        const bool tr = trace && trace->on && trace->launched == 0;
        if (tr) {
            trace->plan = trace->now();
            cudaEventCreate(&ev0);
            cudaEventCreate(&ev1);
            cudaEventRecord(ev0, s_);
        }
        const unsigned long long* gate = gate_;
        // Queue-order upload only when the launch has a dynamic queue (more than one round): in a one-round launch every
        // warp waits for its first stream anyway, so the blob-order upload in few large copies finishes first (C2, 4 096 x
        // 32 KiB: one copy per stream costs 22 ms of driver calls and short transfers, 55.8 vs 33.9 ms per call).
        if (gate_ && upload_phase == 0 && queue_upload.src && plan.order_big.empty() && !getenv("LZB_GATE_BYTES") &&
            (plan.n_small > plan.n_static || getenv("LZB_GATE_QUEUE"))) {
            rc = arm_gate_queue(ctx, queue_upload.src, queue_upload.lead, queue_upload.in_lo, queue_upload.in_bytes, items,
                                plan.order_small, plan.order_stored, &before_first_kernel);
            if (rc < 0) return rc;
        }
        if (drain && (rc = set_done_words(ctx, n, &gate)) != LZB_RC_OK) return rc;
        rc = launch_plan(ctx, s_, plan, ctx->d_items.as<LzbItem>(), ctx->d_order.as<uint32_t>(), in_, out_,
                         ctx->d_results.as<LzbResult>(), ctx->d_counter.as<unsigned int>(), mirror, gate,
                         &before_first_kernel, &upload_phase);
        if (rc != LZB_RC_OK) return rc;
        if ((rc = upload_rest(&before_first_kernel, &upload_phase)) != LZB_RC_OK) return rc;
        std::vector<uint8_t> copied;
        if (drain) {
            CUDA_TRY(ctx, cudaEventRecord(ctx->kernels_done, s_));
            copied.assign(n, 0);
            std::vector<uint32_t> queue(plan.order_small);
            queue.insert(queue.end(), plan.order_big.begin(), plan.order_big.end());
            rc = drain_finished(ctx, queue, items, out_, hmirror_, copied, 4u * (uint32_t)ctx->sm_count * LZB_MAX_WARPS);
            if (rc != LZB_RC_OK) return rc;
        }
        if (tr) {
            cudaEventRecord(ev1, s_);
            trace->launched = trace->now();
        }
        if (tr) trace->uploaded = trace->now();
        CUDA_TRY(ctx, cudaMemcpyAsync(results, ctx->d_results.p, n * sizeof(LzbResult), cudaMemcpyDeviceToHost, s_));
        CUDA_TRY(ctx, cudaStreamSynchronize(s_));
        if (drain) {  // what carried no done word (streams of the stored-chunk copy kernel), then wait for the copy engine
            for (uint32_t i = 0; i < n; i++)
                if (!copied[i] && results[i].out_len && !(items[i].flags & LZB_ITEM_F_OUT_SCRATCH))
                    CUDA_TRY(ctx, cudaMemcpyAsync(hmirror_ + items[i].out_off, out_ + items[i].out_off, results[i].out_len,
                                                  cudaMemcpyDefault, ctx->drain_stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->drain_stream));
        }
        if (gate_ && upload_phase == 2 && !gate_opened_) {
            // later launches of this call (second passes, chained .xz stages) use other queues: the upload has been
            // consumed by the kernel that just finished; open the gate for good
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
            ctx->h_marks[7] = ~0ull;
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_gate.p, &ctx->h_marks[7], 8, cudaMemcpyHostToDevice, s_));
            gate_opened_ = true;
        }
        if (tr) {
            trace->synced = trace->now();
            float ms = 0;
            cudaEventElapsedTime(&ms, ev0, ev1);
            trace->kernel_ms = ms;
            cudaEventDestroy(ev0);
            cudaEventDestroy(ev1);
        }
        return LZB_RC_OK;
    }
    lzb_ctx* ctx_;
    cudaStream_t s_;
    const uint8_t* in_;
    uint8_t* out_;
    uint8_t* hmirror_;
    const unsigned long long* gate_;
    bool unmirrored_ = false;
    bool gate_opened_ = false;
    std::vector<void*> scratch_;
};

struct CopyJoin {  // no exit path may leave copies in flight into ctx->d_in
    cudaStream_t s;
    ~CopyJoin() {
        if (s) cudaStreamSynchronize(s);
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int lzb_abi_version(void) { return LZB_ABI_VERSION; }

extern "C" int lzb_create(lzb_ctx** out, int device) {
    if (!out) return LZB_RC_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return LZB_RC_NO_DEVICE;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return LZB_RC_NO_DEVICE;
    if (device >= count) return LZB_RC_BAD_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return LZB_RC_NO_DEVICE;
    lzb_ctx* ctx = new lzb_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->drain_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->kernels_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->upload_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->gate_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->h_marks, (8 + LZB_GATE_MARKS + 8) * sizeof(unsigned long long), cudaHostAllocDefault) !=
            cudaSuccess) {
        lzb_destroy(ctx);
        return LZB_RC_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    for (k1_t k : k1_variants)
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin) != cudaSuccess) {
            lzb_destroy(ctx);
            return LZB_RC_CUDA;
        }
    *out = ctx;
    return LZB_RC_OK;
}

extern "C" void lzb_destroy(lzb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->drain_stream) cudaStreamSynchronize(ctx->drain_stream);
    if (ctx->oneshot) {
        lzb_batch_destroy(ctx->oneshot);
        ctx->oneshot = nullptr;
    }
    if (ctx->peer_batch) {
        lzb_batch_destroy(ctx->peer_batch);
        ctx->peer_batch = nullptr;
    }
    DevBuf* bufs[] = {&ctx->d_in, &ctx->d_out, &ctx->d_items, &ctx->d_results, &ctx->d_order, &ctx->d_counter, &ctx->d_scan,
                      &ctx->d_off, &ctx->d_crc_ranges, &ctx->d_crc_segmap, &ctx->d_crc_part32, &ctx->d_crc_part64,
                      &ctx->d_crc_out32, &ctx->d_crc_out64, &ctx->d_litws, &ctx->d_matchws, &ctx->d_gate,
                      &ctx->d_enc_items, &ctx->d_enc_results, &ctx->d_enc_pieces, &ctx->d_cap};
    for (DevBuf* b : bufs) b->release();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->drain_stream) cudaStreamDestroy(ctx->drain_stream);
    if (ctx->gate_ready) cudaEventDestroy(ctx->gate_ready);
    if (ctx->kernels_done) cudaEventDestroy(ctx->kernels_done);
    if (ctx->upload_done) cudaEventDestroy(ctx->upload_done);
    if (ctx->h_marks) cudaFreeHost(ctx->h_marks);
    if (ctx->h_done) cudaFreeHost(ctx->h_done);
    delete ctx;
}

extern "C" const char* lzb_last_error(const lzb_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

extern "C" int lzb_scan(lzb_ctx* ctx, int fmt, const lzb_options* opt, const uint8_t* in, const uint64_t* in_off, uint32_t n,
                        uint64_t* capacity) {
    if (!ctx || !in_off || !capacity || (n && !in) || fmt < 0 || fmt > 2) return LZB_RC_BAD_ARG;
    for (uint32_t i = 0; i < n; i++) capacity[i] = lzb::scan_capacity(fmt, opt, in + in_off[i], in_off[i + 1] - in_off[i]);
    return LZB_RC_OK;
}

extern "C" int lzb_decode_batch(lzb_ctx* ctx, int fmt, const lzb_options* opt, const uint8_t* in, const uint64_t* in_off,
                                uint32_t n, uint8_t* out, const uint64_t* out_off, uint64_t* out_len, uint64_t* consumed,
                                lzb_status* st) {
    if (!ctx || !in_off || !out_off || !out_len || !consumed || !st || fmt < 0 || fmt > 2) return LZB_RC_BAD_ARG;
    if (n == 0) return LZB_RC_OK;
    if (!in || !out) return LZB_RC_BAD_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    Trace trace;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t in_lo = in_off[0], in_hi = in_off[n], out_lo = out_off[0], out_hi = out_off[n];
    // device copies keep the low 4 address bits of the host offsets so that (base + offset) stays 16-byte congruent
    CUDA_TRY(ctx, ctx->d_in.ensure((in_hi - in_lo) + 64));
    CUDA_TRY(ctx, ctx->d_out.ensure((out_hi - out_lo) + 64));
    uint8_t* d_in0 = ctx->d_in.as<uint8_t>() + (in_lo & 15);
    uint8_t* d_out0 = ctx->d_out.as<uint8_t>() + (out_lo & 15);
    // A pinned (page-locked, device-mapped) output buffer lets K1 stream finished pages straight to the host while it
    // decodes; a pageable buffer gets one device-to-host copy after the kernel.
    uint8_t* host_mirror = nullptr;
    {
        cudaPointerAttributes attr;
        if (!getenv("LZB_NO_MIRROR") && cudaPointerGetAttributes(&attr, out) == cudaSuccess &&
            attr.type == cudaMemoryTypeHost && attr.devicePointer &&
            ((uintptr_t)attr.devicePointer & 15) == 0)
            host_mirror = (uint8_t*)attr.devicePointer;
        else
            cudaGetLastError();  // clear the "not registered" error of a pageable pointer
    }
    // Input upload.  With a pinned output (mirror kernels) and a blob worth splitting, K1 is launched first and the blob
    // follows in chunks on the copy stream: every warp waits only for its own stream's bytes (input_arrived() in
    // lzb_kernels.cu), so the upload overlaps the decode.  Otherwise: one copy, in stream order before the kernel.
    const unsigned long long* d_gate = nullptr;
    std::function<int(int)> upload;
    CopyJoin join{nullptr};
    const uint64_t in_bytes = in_hi - in_lo, lead = in_lo & 15;
    if (host_mirror && in_bytes >= 2 * LZB_GATE_CHUNK && !getenv("LZB_NO_GATE")) {
        join.s = ctx->copy_stream;
        if (int rc = arm_gate(ctx, in + in_lo, lead, in_lo, in_bytes, &upload, &d_gate)) return rc;
    } else if (in_bytes) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_in0, in + in_lo, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    CudaExecutor ex(ctx, ctx->stream, d_in0 - in_lo, d_out0 - out_lo, host_mirror, d_gate);
    ex.before_first_kernel = upload;
    ex.copy_stream = upload ? ctx->copy_stream : nullptr;
    bool in_pinned = false;  // one copy per stream only pays from page-locked memory (pageable copies are staged)
    {
        cudaPointerAttributes ia;
        if (cudaPointerGetAttributes(&ia, in) == cudaSuccess)
            in_pinned = ia.type == cudaMemoryTypeHost || ia.type == cudaMemoryTypeDevice || ia.type == cudaMemoryTypeManaged;
        else
            cudaGetLastError();
    }
    if (upload && in_pinned && fmt != LZB_FMT_XZ) {  // .xz: blocks are re-planned from the bytes, the whole blob must arrive
        ex.queue_upload.src = in + in_lo;
        ex.queue_upload.lead = lead;
        ex.queue_upload.in_lo = in_lo;
        ex.queue_upload.in_bytes = in_bytes;
    }
    ex.trace = &trace;
    const double t_enq = trace.now();
    std::vector<lzb::StreamOut> outs(n);
    int rc = lzb::decode_batch(ex, fmt, opt, in, in_off, n, out_off, outs.data());  // planning reads the host copy
    if (rc != LZB_RC_OK) return rc;
    if (out_hi > out_lo && !ex.all_mirrored())
        CUDA_TRY(ctx, cudaMemcpyAsync(out + out_lo, d_out0, out_hi - out_lo, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < n; i++) {
        out_len[i] = outs[i].out_len;
        consumed[i] = outs[i].consumed;
        st[i] = outs[i].st;
    }
    if (trace.on)
        fprintf(stderr, "lzb_trace n=%u in=%llu gated=%d setup=%.3f planned=%.3f launched=%.3f uploaded=%.3f synced=%.3f "
                        "end=%.3f kernel_in_situ=%.3f ms\n", n, (unsigned long long)in_bytes, d_gate != nullptr, t_enq,
                trace.plan, trace.launched, trace.uploaded, trace.synced, trace.now(), trace.kernel_ms);
    return LZB_RC_OK;
}

// ---- device-resident batches ----
struct lzb_batch {
    lzb_ctx* ctx = nullptr;
    uint32_t n = 0;
    int fmt = 0;
    const uint8_t* d_in = nullptr;
    uint8_t* d_out = nullptr;
    DevBuf d_items, d_results, d_order, d_counter, d_scan, d_off, d_matchws, d_litws;
    std::vector<LzbItem> items;  // host copy (hdr_len, preset info)
    bool allow_incomplete = false;
    DecodePlan plan;
    uint32_t lclp_hint = 0;
    uint64_t stored_bytes = 0;
    bool host_io = false;
    DevBuf d_redo_items, d_redo_results, d_redo_order, d_redo_counter;  // second pass of lzb_batch_collect (rare)
};

// Fills `b` (a fresh batch, or one whose device buffers are being reused) for one batch of streams.
// dev_offsets: in_off / out_off are device arrays.  mirror_base: see lzb_scan_kernel (0 = no mirror; with a mirror every
// stream goes through K1: the stored-chunk copy kernel does not stream pages).  scan_in: blob the framing scan reads
// (nullptr = d_in; lzb_decode_batch_peer scans the remote copy while d_in is still being filled).
static int batch_prepare_into(lzb_ctx* ctx, int fmt, const lzb_options* opt, const uint8_t* d_in, const uint64_t* in_off,
                              uint32_t n, uint8_t* d_out, const uint64_t* out_off, lzb_batch* b, bool dev_offsets = false,
                              uint64_t mirror_base = 0, const uint8_t* scan_in = nullptr, bool take_lock = true,
                              bool host_io = false) {
    static const lzb_options defaults = {0, 0, 0, 0, {0, 0, 0, 0}, 0, 0};
    if (!ctx || !in_off || !out_off || (fmt != LZB_FMT_LZMA && fmt != LZB_FMT_LZMA2) || n == 0 || !d_in || !d_out)
        return LZB_RC_BAD_ARG;
    if (((uintptr_t)d_in & 15) || ((uintptr_t)d_out & 15)) return LZB_RC_BAD_ARG;
    std::unique_lock<std::mutex> lock(ctx->mu, std::defer_lock);
    if (take_lock) lock.lock();
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    b->ctx = ctx;
    b->n = n;
    b->fmt = fmt;
    b->d_in = d_in;
    b->d_out = d_out;
    b->allow_incomplete = opt && opt->allow_incomplete;
    cudaStream_t s = ctx->stream;
    auto fail = [&](int code) { return code; };
#define B_TRY(call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            snprintf(ctx->err, sizeof(ctx->err), "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(e_ == cudaErrorMemoryAllocation ? LZB_RC_OOM : LZB_RC_CUDA);                  \
        }                                                                                             \
    } while (0)
    B_TRY(b->d_items.ensure(n * sizeof(LzbItem)));
    B_TRY(b->d_results.ensure(n * sizeof(LzbResult)));
    B_TRY(b->d_counter.ensure(64));
    B_TRY(b->d_scan.ensure(n * sizeof(LzbScan)));
    B_TRY(b->d_off.ensure(2 * (size_t)(n + 1) * 8));
    uint64_t* d_in_off = b->d_off.as<uint64_t>();
    uint64_t* d_out_off = d_in_off + (n + 1);
    const cudaMemcpyKind off_kind = dev_offsets ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    B_TRY(cudaMemcpyAsync(d_in_off, in_off, (n + 1) * 8, off_kind, s));
    B_TRY(cudaMemcpyAsync(d_out_off, out_off, (n + 1) * 8, off_kind, s));
    lzb_scan_kernel<<<(n + 127) / 128, 128, 0, s>>>(fmt, opt ? *opt : defaults, scan_in ? scan_in : d_in, d_in_off,
                                                    d_out_off, n, b->d_items.as<LzbItem>(), b->d_scan.as<LzbScan>(),
                                                    mirror_base, nullptr);
    B_TRY(cudaGetLastError());
    std::vector<LzbScan> scan(n);
    b->items.resize(n);
    B_TRY(cudaMemcpyAsync(scan.data(), b->d_scan.p, n * sizeof(LzbScan), cudaMemcpyDeviceToHost, s));
    B_TRY(cudaMemcpyAsync(b->items.data(), b->d_items.p, n * sizeof(LzbItem), cudaMemcpyDeviceToHost, s));
    B_TRY(cudaStreamSynchronize(s));
    uint32_t lclp = 0;
    uint64_t stored_bytes = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (b->items[i].kind == LZB_ITEM_LZMA2) lclp = std::max<uint32_t>(lclp, scan[i].max_lclp);
        stored_bytes += scan[i].stored;
    }
    b->lclp_hint = lclp;
    b->stored_bytes = stored_bytes;
    make_plan(ctx, b->items.data(), n, lclp, stored_bytes, &b->plan, /*route_stored=*/mirror_base == 0, host_io);
    b->host_io = host_io;
    if (upload_order(ctx, s, b->plan, b->d_order) != LZB_RC_OK) return fail(LZB_RC_CUDA);
    B_TRY(cudaStreamSynchronize(s));
    return LZB_RC_OK;
#undef B_TRY
}

// Second pass over the few streams the first launch could not finish (same rule as CudaExecutor::decode on the host
// path): LZMA2 streams whose framing scan under-estimated lc+lp -- the scan skips `packed` bytes per chunk while the
// decoder continues where the range coder stopped (SURVEY 3.5 leniency (d)), so a later chunk header it never saw can ask
// for more -- are rerun with tables sized for lc+lp = 4; streams whose input missed the gate's timeout are rerun once the
// copy stream has drained.  `res` is patched in place.
static int batch_redo(lzb_batch* b, cudaStream_t s, std::vector<LzbResult>& res, bool mirror) {
    lzb_ctx* ctx = b->ctx;
    std::vector<uint32_t> redo;
    for (uint32_t i = 0; i < b->n; i++) {
        const bool lclp = b->items[i].kind == LZB_ITEM_LZMA2 && res[i].code == LZB_E_UNSUPPORTED && res[i].a0 <= 4 &&
                          res[i].a1 < 4;
        if (lclp || res[i].code == LZB_E_INPUT_TIMEOUT) redo.push_back(i);
    }
    if (redo.empty()) return LZB_RC_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    const uint32_t m = (uint32_t)redo.size();
    std::vector<LzbItem> sub(m);
    for (uint32_t k = 0; k < m; k++) sub[k] = b->items[redo[k]];
    DecodePlan plan;
    make_plan(ctx, sub.data(), m, 4, b->stored_bytes, &plan, /*route_stored=*/!mirror, b->host_io);
    CUDA_TRY(ctx, b->d_redo_items.ensure(m * sizeof(LzbItem)));
    CUDA_TRY(ctx, b->d_redo_results.ensure(m * sizeof(LzbResult)));
    CUDA_TRY(ctx, b->d_redo_counter.ensure(64));
    CUDA_TRY(ctx, cudaMemcpyAsync(b->d_redo_items.p, sub.data(), m * sizeof(LzbItem), cudaMemcpyHostToDevice, s));
    if (int rc = upload_order(ctx, s, plan, b->d_redo_order)) return rc;
    if (int rc = launch_plan(ctx, s, plan, b->d_redo_items.as<LzbItem>(), b->d_redo_order.as<uint32_t>(), b->d_in, b->d_out,
                             b->d_redo_results.as<LzbResult>(), b->d_redo_counter.as<unsigned int>(), mirror, nullptr,
                             nullptr, nullptr, &b->d_matchws, &b->d_litws))
        return rc;
    std::vector<LzbResult> subres(m);
    CUDA_TRY(ctx, cudaMemcpyAsync(subres.data(), b->d_redo_results.p, m * sizeof(LzbResult), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    for (uint32_t k = 0; k < m; k++) res[redo[k]] = subres[k];
    return LZB_RC_OK;
}

extern "C" int lzb_batch_prepare(lzb_ctx* ctx, int fmt, const lzb_options* opt, const uint8_t* d_in, const uint64_t* in_off,
                                 uint32_t n, uint8_t* d_out, const uint64_t* out_off, lzb_batch** out) {
    if (!out) return LZB_RC_BAD_ARG;
    *out = nullptr;
    lzb_batch* b = new lzb_batch();
    int rc = batch_prepare_into(ctx, fmt, opt, d_in, in_off, n, d_out, out_off, b);
    if (rc != LZB_RC_OK) {
        if (ctx) {
            b->ctx = ctx;
            lzb_batch_destroy(b);
        } else {
            delete b;
        }
        return rc;
    }
    *out = b;
    return LZB_RC_OK;
}

extern "C" int lzb_batch_launch(lzb_batch* b, void* cuda_stream) {
    if (!b) return LZB_RC_BAD_ARG;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : b->ctx->stream;
    // a caller driving several GPUs from one thread may have another device current
    if (cudaSetDevice(b->ctx->device) != cudaSuccess) return LZB_RC_CUDA;
    return launch_plan(b->ctx, s, b->plan, b->d_items.as<LzbItem>(), b->d_order.as<uint32_t>(), b->d_in, b->d_out,
                       b->d_results.as<LzbResult>(), b->d_counter.as<unsigned int>(), false, nullptr, nullptr, nullptr,
                       &b->d_matchws, &b->d_litws);
}

extern "C" int lzb_batch_kernels_per_launch(const lzb_batch* b) {
    return b ? (int)!b->plan.order_small.empty() + (int)!b->plan.order_big.empty() + (int)!b->plan.order_stored.empty() : 0;
}

extern "C" const char* lzb_batch_kernel_name(const lzb_batch* b) {
    if (!b) return "";
    if (!b->plan.order_small.empty()) return k1_names[k1_variant(b->plan, false)];
    if (!b->plan.order_big.empty()) return k1_names[12];
    return b->plan.order_stored.empty() ? "" : "lzb_stored_decode_kernel";
}

extern "C" int lzb_batch_collect(lzb_batch* b, void* cuda_stream, uint64_t* out_len, uint64_t* consumed, lzb_status* st) {
    if (!b) return LZB_RC_BAD_ARG;
    lzb_ctx* ctx = b->ctx;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<LzbResult> res(b->n);
    CUDA_TRY(ctx, cudaMemcpyAsync(res.data(), b->d_results.p, b->n * sizeof(LzbResult), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    if (int rc = batch_redo(b, s, res, false)) return rc;
    lzb_options o = {};
    o.allow_incomplete = b->allow_incomplete;
    for (uint32_t i = 0; i < b->n; i++) {
        if (lzb::lenient_eof(b->fmt, &o, &res[i])) res[i].code = LZB_OK, res[i].sink_len = res[i].out_len;
        if (st) lzb::status_from_result(res[i], &st[i]);
        if (out_len) out_len[i] = res[i].sink_len;
        if (consumed) consumed[i] = b->items[i].hdr_len + res[i].consumed;
    }
    return LZB_RC_OK;
}

extern "C" void lzb_batch_destroy(lzb_batch* b) {
    if (!b) return;
    if (b->ctx) cudaSetDevice(b->ctx->device);
    DevBuf* bufs[] = {&b->d_items, &b->d_results, &b->d_order, &b->d_counter, &b->d_scan, &b->d_off, &b->d_matchws,
                      &b->d_litws, &b->d_redo_items, &b->d_redo_results, &b->d_redo_order, &b->d_redo_counter};
    for (DevBuf* x : bufs) x->release();
    delete b;
}

extern "C" int lzb_decode_batch_device(lzb_ctx* ctx, int fmt, const lzb_options* opt, const uint8_t* d_in,
                                       const uint64_t* in_off, uint32_t n, uint8_t* d_out, const uint64_t* out_off,
                                       uint64_t* out_len, uint64_t* consumed, lzb_status* st, void* cuda_stream) {
    if (n == 0) return LZB_RC_OK;
    if (!ctx) return LZB_RC_BAD_ARG;
    std::lock_guard<std::mutex> one(ctx->oneshot_mu);
    if (!ctx->oneshot) {
        ctx->oneshot = new lzb_batch();
        ctx->oneshot->ctx = ctx;
    }
    lzb_batch* b = ctx->oneshot;
    int rc = batch_prepare_into(ctx, fmt, opt, d_in, in_off, n, d_out, out_off, b);
    if (rc != LZB_RC_OK) return rc;
    if (cuda_stream) cudaStreamSynchronize(ctx->stream);  // prepare ran on the ctx stream
    rc = lzb_batch_launch(b, cuda_stream);
    if (rc == LZB_RC_OK) rc = lzb_batch_collect(b, cuda_stream, out_len, consumed, st);
    return rc;
}

extern "C" int lzb_batch_prepare_device(lzb_ctx* ctx, int fmt, const lzb_options* opt, const uint8_t* d_in,
                                        const uint64_t* d_in_off, uint32_t n, uint8_t* d_out, const uint64_t* d_out_off,
                                        lzb_batch** out) {
    if (!out) return LZB_RC_BAD_ARG;
    *out = nullptr;
    lzb_batch* b = new lzb_batch();
    int rc = batch_prepare_into(ctx, fmt, opt, d_in, d_in_off, n, d_out, d_out_off, b, /*dev_offsets=*/true);
    if (rc != LZB_RC_OK) {
        if (ctx) {
            b->ctx = ctx;
            lzb_batch_destroy(b);
        } else {
            delete b;
        }
        return rc;
    }
    *out = b;
    return LZB_RC_OK;
}

// Sizes + output layout of a device-resident batch, all on the device (K2 + lzb_layout_kernel).
extern "C" int lzb_scan_device(lzb_ctx* ctx, int fmt, const lzb_options* opt, const uint8_t* d_in, const uint64_t* d_in_off,
                               uint32_t n, uint64_t* d_capacity, uint64_t* d_out_off, uint64_t* total, void* cuda_stream) {
    static const lzb_options defaults = {0, 0, 0, 0, {0, 0, 0, 0}, 0, 0};
    if (!ctx || (fmt != LZB_FMT_LZMA && fmt != LZB_FMT_LZMA2) || !d_in_off || (n && !d_in)) return LZB_RC_BAD_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);  // K2 writes the context's scratch items
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CUDA_TRY(ctx, ctx->d_items.ensure(std::max<size_t>(n, 1) * sizeof(LzbItem)));
    CUDA_TRY(ctx, ctx->d_scan.ensure(std::max<size_t>(n, 1) * sizeof(LzbScan)));
    uint64_t* cap = d_capacity;
    if (!cap) {
        CUDA_TRY(ctx, ctx->d_cap.ensure(std::max<size_t>(n, 1) * 8));
        cap = ctx->d_cap.as<uint64_t>();
    }
    if (n) {
        lzb_scan_kernel<<<(n + 127) / 128, 128, 0, s>>>(fmt, opt ? *opt : defaults, d_in, d_in_off, nullptr, n,
                                                        ctx->d_items.as<LzbItem>(), ctx->d_scan.as<LzbScan>(), 0, cap);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (d_out_off) {
        lzb_layout_kernel<<<1, 1024, 0, s>>>(cap, n, d_out_off);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (total) {
        *total = 0;
        if (d_out_off) {
            CUDA_TRY(ctx, cudaMemcpyAsync(total, d_out_off + n, 8, cudaMemcpyDeviceToHost, s));
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(s));
    }
    return LZB_RC_OK;
}

// ---- CUDA IPC: a blob in one process's HBM, mapped into the other ranks of the node (NVLink peer access) ----
extern "C" int lzb_ipc_export(lzb_ctx* ctx, const void* d_ptr, uint64_t bytes, lzb_ipc_handle* h) {
    if (!ctx || !d_ptr || !h) return LZB_RC_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(h->handle), "lzb_ipc_handle::handle is a cudaIpcMemHandle_t");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // the handle names the whole allocation; find its base through the driver entry point (no libcuda link dependency)
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(ctx, cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    unsigned long long base = 0;
    size_t size = 0;
    if (!fn || ((range_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)d_ptr) != 0) {
        snprintf(ctx->err, sizeof(ctx->err), "cuMemGetAddressRange failed for %p", d_ptr);
        return LZB_RC_CUDA;
    }
    cudaIpcMemHandle_t mh;
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&mh, (void*)(uintptr_t)base));
    memcpy(h->handle, &mh, sizeof mh);
    h->offset = (uint64_t)(uintptr_t)d_ptr - base;
    h->bytes = bytes;
    return LZB_RC_OK;
}

extern "C" int lzb_ipc_open(lzb_ctx* ctx, const lzb_ipc_handle* h, void** d_ptr) {
    if (!ctx || !h || !d_ptr) return LZB_RC_BAD_ARG;
    *d_ptr = nullptr;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t mh;
    memcpy(&mh, h->handle, sizeof mh);
    void* base = nullptr;
    CUDA_TRY(ctx, cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess));
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        ctx->ipc_open.push_back({(uint8_t*)base + h->offset, base});
    }
    *d_ptr = (uint8_t*)base + h->offset;
    return LZB_RC_OK;
}

extern "C" int lzb_ipc_close(lzb_ctx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return LZB_RC_BAD_ARG;
    void* base = nullptr;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        for (size_t i = 0; i < ctx->ipc_open.size(); i++)
            if (ctx->ipc_open[i].first == d_ptr) {
                base = ctx->ipc_open[i].second;
                ctx->ipc_open.erase(ctx->ipc_open.begin() + i);
                break;
            }
    }
    if (!base) return LZB_RC_BAD_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaIpcCloseMemHandle(base));
    return LZB_RC_OK;
}

// Scatter + decode + gather for one rank's stream range as ONE launch: see include/lzma_b200.h.
extern "C" int lzb_decode_batch_peer(lzb_ctx* ctx, int fmt, const lzb_options* opt, const uint8_t* src_in,
                                     const uint64_t* in_off, uint32_t n, uint8_t* dst_out, const uint64_t* out_off,
                                     uint64_t* out_len, uint64_t* consumed, lzb_status* st) {
    if (!ctx || !in_off || !out_off || (fmt != LZB_FMT_LZMA && fmt != LZB_FMT_LZMA2)) return LZB_RC_BAD_ARG;
    if (n == 0) return LZB_RC_OK;
    if (!src_in || !dst_out || ((uintptr_t)dst_out & 15)) return LZB_RC_BAD_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    Trace trace;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t in_lo = in_off[0], in_hi = in_off[n], out_lo = out_off[0], out_hi = out_off[n];
    const uint64_t in_bytes = in_hi - in_lo;
    // local staging keeps the 16-byte congruence of the remote offsets (and the 4-byte congruence K2 / K1 rely on)
    const uint64_t lead = ((uintptr_t)src_in + in_lo) & 15;
    CUDA_TRY(ctx, ctx->d_in.ensure(in_bytes + 256));
    CUDA_TRY(ctx, ctx->d_out.ensure((out_hi - out_lo) + 64));
    uint8_t* d_in0 = ctx->d_in.as<uint8_t>() + lead;
    uint8_t* d_out0 = ctx->d_out.as<uint8_t>() + (out_lo & 15);
    if (!ctx->peer_batch) {
        ctx->peer_batch = new lzb_batch();
        ctx->peer_batch->ctx = ctx;
    }
    lzb_batch* b = ctx->peer_batch;
    // K1 addresses its blobs as base + offset: hand it bases that put offset in_lo / out_lo on the local copies.  Their low
    // four bits: d_in0 - in_lo == src_in (mod 16) by the choice of `lead`; K2 reads the remote blob in place.
    const uint8_t* in_base = d_in0 - in_lo;
    uint8_t* out_base = d_out0 - out_lo;
    const uint8_t* scan_base = src_in;
    // batch_prepare_into wants 16-byte aligned bases: fold the misalignment of src_in into shifted offset arrays
    std::vector<uint64_t> in_adj;
    const uint64_t mis = (uintptr_t)src_in & 15;
    if (mis) {
        in_adj.assign(in_off, in_off + n + 1);
        for (auto& v : in_adj) v += mis;
        in_off = in_adj.data();
        scan_base -= mis;
        in_base -= mis;
    }
    // same two forms as the host API (CudaExecutor::run): page stores by K1 for up to two rounds of streams, done words +
    // copy engine (peer-to-peer over NVLink) beyond
    const char* dr = getenv("LZB_DRAIN_ROUNDS");
    const bool drain = !getenv("LZB_NO_DRAIN") && (double)n > (dr ? atof(dr) : 1.0) * (double)ctx->sm_count * LZB_MAX_WARPS;
    int rc = batch_prepare_into(ctx, fmt, opt, in_base, in_off, n, out_base, out_off, b, false,
                                drain ? 0 : (uint64_t)(uintptr_t)dst_out, scan_base, /*take_lock=*/false, /*host_io=*/true);
    if (rc != LZB_RC_OK) return rc;
    bool all_mirrored = true;
    for (uint32_t i = 0; i < n; i++)
        if (b->items[i].kind != LZB_ITEM_PRESET && (out_off[i] & 15)) all_mirrored = false;
    if (drain) all_mirrored = true;  // the copy engine has no alignment needs
    trace.plan = trace.now();
    // the input follows behind the gate; small shards are not worth the chunking
    const unsigned long long* d_gate = nullptr;
    std::function<int(int)> upload;
    int upload_phase = 0;
    CopyJoin join{nullptr};
    if (in_bytes >= 2 * LZB_GATE_CHUNK && !getenv("LZB_NO_GATE")) {
        join.s = ctx->copy_stream;
        if ((rc = arm_gate(ctx, src_in + (in_off[0] - mis), lead, in_off[0], in_bytes, &upload, &d_gate))) return rc;
        if (b->plan.order_big.empty() && !getenv("LZB_GATE_BYTES") &&
            (rc = arm_gate_queue(ctx, src_in + (in_off[0] - mis), lead, in_off[0], in_bytes, b->items.data(), b->plan.order_small,
                                 b->plan.order_stored, &upload)) < 0)
            return rc;
    } else {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_in0, src_in + (in_off[0] - mis), in_bytes, cudaMemcpyDefault, ctx->stream));
    }
    cudaStream_t s = ctx->stream;
    if (drain && (rc = set_done_words(ctx, n, &d_gate)) != LZB_RC_OK) return rc;
    rc = launch_plan(ctx, s, b->plan, b->d_items.as<LzbItem>(), b->d_order.as<uint32_t>(), in_base, out_base,
                     b->d_results.as<LzbResult>(), b->d_counter.as<unsigned int>(), /*mirror=*/!drain, d_gate, &upload,
                     &upload_phase, &b->d_matchws, &b->d_litws);
    if (rc != LZB_RC_OK) return rc;
    if ((rc = upload_rest(&upload, &upload_phase)) != LZB_RC_OK) return rc;
    trace.launched = trace.now();
    std::vector<uint8_t> copied;
    if (drain) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->kernels_done, s));
        copied.assign(n, 0);
        std::vector<uint32_t> queue(b->plan.order_small);
        queue.insert(queue.end(), b->plan.order_big.begin(), b->plan.order_big.end());
        rc = drain_finished(ctx, queue, b->items.data(), out_base, dst_out, copied, 4u * (uint32_t)ctx->sm_count * LZB_MAX_WARPS);
        if (rc != LZB_RC_OK) return rc;
    }
    std::vector<LzbResult> res(n);
    CUDA_TRY(ctx, cudaMemcpyAsync(res.data(), b->d_results.p, n * sizeof(LzbResult), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    const std::vector<LzbResult> first_pass(res);
    if ((rc = batch_redo(b, s, res, !drain))) return rc;
    if (drain) {  // streams without a done word (stored-chunk kernel) and streams the second pass decoded again
        for (uint32_t i = 0; i < n; i++) {
            const bool redone = first_pass[i].code != res[i].code || first_pass[i].out_len != res[i].out_len;
            if ((!copied[i] || redone) && res[i].out_len)
                CUDA_TRY(ctx, cudaMemcpyAsync(dst_out + b->items[i].out_off, out_base + b->items[i].out_off, res[i].out_len,
                                              cudaMemcpyDefault, ctx->drain_stream));
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->drain_stream));
    }
    if (out_hi > out_lo && !all_mirrored)  // some stream's region was not 16-byte aligned: plain copy of the whole range
        CUDA_TRY(ctx, cudaMemcpyAsync(dst_out + out_lo, d_out0, out_hi - out_lo, cudaMemcpyDefault, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    lzb_options o = {};
    o.allow_incomplete = b->allow_incomplete;
    for (uint32_t i = 0; i < n; i++) {
        if (lzb::lenient_eof(fmt, &o, &res[i])) res[i].code = LZB_OK, res[i].sink_len = res[i].out_len;
        if (st) lzb::status_from_result(res[i], &st[i]);
        if (out_len) out_len[i] = res[i].sink_len;
        if (consumed) consumed[i] = b->items[i].hdr_len + res[i].consumed;
    }
    if (trace.on)
        fprintf(stderr, "lzb_trace peer n=%u in=%llu gated=%d planned=%.3f launched=%.3f end=%.3f ms\n", n,
                (unsigned long long)in_bytes, d_gate != nullptr, trace.plan, trace.launched, trace.now());
    return LZB_RC_OK;
}

// ---- one process, several devices ----
struct lzb_multi {
    std::vector<lzb_ctx*> ctxs;
    char err[400] = {0};
};

extern "C" int lzb_create_multi(lzb_multi** out, const int* dev_ids, int n_dev) {
    if (!out) return LZB_RC_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return LZB_RC_NO_DEVICE;
    std::vector<int> ids;
    if (!dev_ids || n_dev <= 0)
        for (int d = 0; d < count; d++) ids.push_back(d);
    else
        ids.assign(dev_ids, dev_ids + n_dev);
    lzb_multi* m = new lzb_multi();
    for (int d : ids) {
        lzb_ctx* c = nullptr;
        int rc = (d < 0 || d >= count) ? LZB_RC_BAD_ARG : lzb_create(&c, d);
        if (rc != LZB_RC_OK) {
            lzb_destroy_multi(m);
            return rc;
        }
        m->ctxs.push_back(c);
    }
    *out = m;
    return LZB_RC_OK;
}

extern "C" void lzb_destroy_multi(lzb_multi* m) {
    if (!m) return;
    for (lzb_ctx* c : m->ctxs) lzb_destroy(c);
    delete m;
}

extern "C" int lzb_multi_device_count(const lzb_multi* m) { return m ? (int)m->ctxs.size() : 0; }
extern "C" lzb_ctx* lzb_multi_ctx(lzb_multi* m, int k) { return m && k >= 0 && k < (int)m->ctxs.size() ? m->ctxs[k] : nullptr; }
extern "C" const char* lzb_multi_last_error(const lzb_multi* m) { return m ? m->err : "null multi"; }

extern "C" int lzb_decode_batch_multi(lzb_multi* m, int fmt, const lzb_options* opt, const uint8_t* in, const uint64_t* in_off,
                                      uint32_t n, uint8_t* out, const uint64_t* out_off, uint64_t* out_len,
                                      uint64_t* consumed, lzb_status* st, uint32_t* split) {
    if (!m || m->ctxs.empty() || !in_off || !out_off || !out_len || !consumed || !st) return LZB_RC_BAD_ARG;
    const uint32_t nd = (uint32_t)m->ctxs.size();
    // contiguous stream ranges with equal shares of the compressed bytes (decode time is proportional to decisions, for
    // which the compressed length is the cheap proxy); contiguous, so every device's call is a plain sub-range of the
    // caller's arrays and its uploads / page stores touch one slice of the caller's buffers
    std::vector<uint32_t> cut(nd + 1, n);
    cut[0] = 0;
    const uint64_t base = n ? in_off[0] : 0, total = n ? in_off[n] - in_off[0] : 0;
    for (uint32_t k = 1; k < nd; k++) {
        const uint64_t target = base + (total / nd) * k + (total % nd) * k / nd;
        uint32_t c = (uint32_t)(std::lower_bound(in_off, in_off + n + 1, target) - in_off);
        cut[k] = std::min(std::max(c, cut[k - 1]), n);
    }
    if (split) memcpy(split, cut.data(), (nd + 1) * sizeof(uint32_t));
    std::vector<int> rcs(nd, LZB_RC_OK);
    std::vector<std::thread> th;
    for (uint32_t k = 0; k < nd; k++) {
        const uint32_t lo = cut[k], cnt = cut[k + 1] - cut[k];
        if (!cnt) continue;
        th.emplace_back([=, &rcs]() {
            rcs[k] = lzb_decode_batch(m->ctxs[k], fmt, opt, in, in_off + lo, cnt, out, out_off + lo, out_len + lo,
                                      consumed + lo, st + lo);
        });
    }
    for (auto& t : th) t.join();
    for (uint32_t k = 0; k < nd; k++)
        if (rcs[k] != LZB_RC_OK) {
            snprintf(m->err, sizeof m->err, "device %d: %s", m->ctxs[k]->device, m->ctxs[k]->err);
            return rcs[k];
        }
    return LZB_RC_OK;
}

extern "C" int lzb_decompress_alloc(lzb_ctx* ctx, int fmt, const lzb_options* opt, const uint8_t* in, size_t in_len,
                                    uint8_t** out, size_t* out_len, size_t* consumed, lzb_status* st) {
    if (!ctx || !out || !out_len || !consumed || !st || (in_len && !in) || fmt < 0 || fmt > 2) return LZB_RC_BAD_ARG;
    static const uint8_t empty = 0;
    if (!in) in = &empty;
    *out = nullptr;
    *out_len = 0;
    *consumed = 0;
    uint64_t cap = lzb::scan_capacity(fmt, opt, in, in_len);
    const uint64_t limit = 0xFFFFF000ull;
    for (;;) {
        cap = std::min<uint64_t>(cap, limit);
        uint8_t* buf = (uint8_t*)malloc((size_t)cap + 16);
        if (!buf) return LZB_RC_OOM;
        uint64_t in_off[2] = {0, in_len}, out_off[2] = {0, cap}, ol = 0, cons = 0;
        int rc = lzb_decode_batch(ctx, fmt, opt, in, in_off, 1, buf, out_off, &ol, &cons, st);
        if (rc != LZB_RC_OK) {
            free(buf);
            return rc;
        }
        if (st->code == LZB_E_CAPACITY && cap < limit) {  // end-marker .lzma / malformed framing: grow and retry
            free(buf);
            cap = std::max<uint64_t>(cap * 2, st->a0 + 65536);
            continue;
        }
        if (st->code == LZB_E_CAPACITY) {
            st->code = LZB_E_UNSUPPORTED;
            st->kind = LZB_KIND_INTERNAL;
        }
        *out = buf;
        *out_len = (size_t)ol;
        *consumed = (size_t)cons;
        return LZB_RC_OK;
    }
}

extern "C" void lzb_free(void* p) { free(p); }

// ---- decompress::raw decoder objects: DecoderState kept in device memory between calls ----
struct lzb_raw {
    lzb_ctx* ctx = nullptr;
    int fmt = 0;
    uint32_t lc = 0, lp = 0, pb = 0, dict_size = 0, lclp_cap = 0;
    DevBuf state[2];  // [cur]: the committed state; the other one is what the running call works on
    int cur = 0;
    DevBuf d_item, d_result, d_aux;  // d_aux: the one-entry queue {0} and the queue counter of the launch
};

static int raw_write_fresh(lzb_raw* r) {
    lzb_ctx* ctx = r->ctx;
    LzbCarry h = {};
    h.fresh = 1;
    h.lc = r->fmt == LZB_FMT_LZMA ? r->lc : 0;
    h.lp = r->fmt == LZB_FMT_LZMA ? r->lp : 0;
    h.pb = r->fmt == LZB_FMT_LZMA ? r->pb : 0;
    h.lclp_cap = r->lclp_cap;
    CUDA_TRY(ctx, cudaMemcpyAsync(r->state[r->cur].p, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LZB_RC_OK;
}

extern "C" int lzb_raw_create(lzb_ctx* ctx, int fmt, uint32_t lc, uint32_t lp, uint32_t pb, uint32_t dict_size, lzb_raw** out) {
    if (!out) return LZB_RC_BAD_ARG;
    *out = nullptr;
    if (!ctx || (fmt != LZB_FMT_LZMA && fmt != LZB_FMT_LZMA2)) return LZB_RC_BAD_ARG;
    if (fmt == LZB_FMT_LZMA && (lc > 8 || lp > 4 || pb > 4 || dict_size == 0)) return LZB_RC_BAD_ARG;  // lzma.rs:62-66
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    lzb_raw* r = new lzb_raw();
    r->ctx = ctx;
    r->fmt = fmt;
    r->lc = lc;
    r->lp = lp;
    r->pb = pb;
    r->dict_size = dict_size;
    r->lclp_cap = fmt == LZB_FMT_LZMA ? lc + lp : 4;  // LZMA2 chunks may switch to any lc + lp <= 4 (lzma2.rs:170-175)
    const size_t bytes = (size_t)lzb_carry_bytes(r->lclp_cap);
    int rc = LZB_RC_OK;
    if (r->state[0].ensure(bytes) != cudaSuccess || r->state[1].ensure(bytes) != cudaSuccess ||
        r->d_item.ensure(sizeof(LzbItem)) != cudaSuccess || r->d_result.ensure(sizeof(LzbResult)) != cudaSuccess ||
        r->d_aux.ensure(64) != cudaSuccess)
        rc = LZB_RC_OOM;
    if (rc == LZB_RC_OK) rc = raw_write_fresh(r);
    if (rc != LZB_RC_OK) {
        lzb_raw_destroy(r);
        return rc;
    }
    *out = r;
    return LZB_RC_OK;
}

extern "C" int lzb_raw_reset(lzb_raw* r) {
    if (!r) return LZB_RC_BAD_ARG;
    std::lock_guard<std::mutex> lock(r->ctx->mu);
    if (cudaSetDevice(r->ctx->device) != cudaSuccess) return LZB_RC_CUDA;
    return raw_write_fresh(r);
}

extern "C" void lzb_raw_destroy(lzb_raw* r) {
    if (!r) return;
    if (r->ctx) cudaSetDevice(r->ctx->device);
    r->state[0].release();
    r->state[1].release();
    r->d_item.release();
    r->d_result.release();
    r->d_aux.release();
    delete r;
}

extern "C" int lzb_raw_decompress(lzb_raw* r, const lzb_options* opt, const uint8_t* in, size_t in_len, uint8_t** out,
                                  size_t* out_len, size_t* consumed, lzb_status* st) {
    if (!r || !out || !out_len || !consumed || !st || (in_len && !in)) return LZB_RC_BAD_ARG;
    lzb_ctx* ctx = r->ctx;
    *out = nullptr;
    *out_len = *consumed = 0;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    CUDA_TRY(ctx, ctx->d_in.ensure(in_len + 64));
    if (in_len) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_in.p, in, in_len, cudaMemcpyHostToDevice, s));
    const uint64_t limit = 0xFFFFF000ull, bound = (uint64_t)in_len * 16384 + (1u << 20);
    const bool sized = r->fmt == LZB_FMT_LZMA && opt && opt->has_provided;
    uint64_t cap;
    if (r->fmt == LZB_FMT_LZMA2) {
        static const uint8_t empty = 0;
        cap = std::min<uint64_t>(lzb::scan_lzma2(in ? in : &empty, in_len).unpacked, bound) + 16;
    } else {
        cap = sized && opt->provided <= bound ? opt->provided + 288 : (uint64_t)in_len * 8 + 65536;
    }
    const LzbKC kc = LZB_KC_INIT;
    const int smem = ((int)T_LIT * 2 + 15) & ~15;
    const size_t state_bytes = (size_t)lzb_carry_bytes(r->lclp_cap);
    LzbResult res;
    for (;;) {
        cap = std::min(cap, limit);
        CUDA_TRY(ctx, ctx->d_out.ensure(cap + 64));
        const int work = 1 - r->cur;
        CUDA_TRY(ctx, cudaMemcpyAsync(r->state[work].p, r->state[r->cur].p, state_bytes, cudaMemcpyDeviceToDevice, s));
        LzbItem it = {};
        it.in_off = 0;
        it.in_len = in_len;
        it.out_off = 0;
        it.out_cap = cap;
        it.unpacked = sized ? opt->provided : LZB_UNKNOWN_SIZE;
        it.memlimit = opt && opt->has_memlimit ? opt->memlimit : ~0ull;
        it.dict_size = r->dict_size;
        it.kind = r->fmt == LZB_FMT_LZMA ? LZB_ITEM_LZMA : LZB_ITEM_LZMA2;
        it.lc = (uint8_t)r->lc;
        it.lp = (uint8_t)r->lp;
        it.pb = (uint8_t)r->pb;
        it.flags = LZB_ITEM_F_CARRY;
        it.host_out = (uint64_t)(uintptr_t)r->state[work].p;
        CUDA_TRY(ctx, cudaMemcpyAsync(r->d_item.p, &it, sizeof it, cudaMemcpyHostToDevice, s));
        // the lc+lp > 4 kernel with one warp: queue = {0}, counter zeroed (d_aux[0] = order entry, d_aux[8..] = counter)
        CUDA_TRY(ctx, cudaMemsetAsync(r->d_aux.p, 0, 64, s));
        lzb_decode_biglit_kernel<<<1, 32, smem, s>>>(r->d_item.as<LzbItem>(), r->d_aux.as<uint32_t>(), 1u, 0u, ctx->d_in.as<uint8_t>(),
                                                     ctx->d_out.as<uint8_t>(), r->d_result.as<LzbResult>(),
                                                     r->d_aux.as<unsigned int>() + 8, r->lclp_cap, (uint32_t)smem, nullptr, 0ull, kc,
                                                     nullptr);
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaMemcpyAsync(&res, r->d_result.p, sizeof res, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(ctx, cudaStreamSynchronize(s));
        if (res.code == LZB_E_CAPACITY && cap < limit) {  // size unknown up front: same state, larger buffer
            cap = std::max<uint64_t>(cap * 2, res.a0 + 65536);
            continue;
        }
        if (res.code == LZB_OK) r->cur = work;  // commit; a failed call leaves the decoder as it was before the call
        break;
    }
    lzb::status_from_result(res, st);
    if (st->code == LZB_E_CAPACITY) {
        st->code = LZB_E_UNSUPPORTED;
        st->kind = LZB_KIND_INTERNAL;
    }
    uint8_t* buf = (uint8_t*)malloc((size_t)res.sink_len + 16);
    if (!buf) return LZB_RC_OOM;
    if (res.sink_len) {
        CUDA_TRY(ctx, cudaMemcpyAsync(buf, ctx->d_out.p, res.sink_len, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(ctx, cudaStreamSynchronize(s));
    }
    *out = buf;
    *out_len = (size_t)res.sink_len;
    *consumed = (size_t)res.consumed;
    return LZB_RC_OK;
}

extern "C" int lzb_crc_device(lzb_ctx* ctx, const uint8_t* d_data, const uint64_t* off, const uint64_t* len, uint32_t n,
                              uint32_t* crc32, uint64_t* crc64, void* cuda_stream) {
    if (!ctx || !d_data || !off || !len || !crc32 || !crc64) return LZB_RC_BAD_ARG;
    if (n == 0) return LZB_RC_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<lzb::CrcRange> r(n);
    for (uint32_t i = 0; i < n; i++) r[i] = lzb::CrcRange{off[i], len[i]};
    CudaExecutor ex(ctx, cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream, nullptr, const_cast<uint8_t*>(d_data));
    return ex.crc(r.data(), n, crc32, crc64);
}

// ------------------------------------------------------------------------------------------------
// compress side (lzb_encode_kernels.cu)
// ------------------------------------------------------------------------------------------------
extern "C" uint64_t lzb_encode_bound(int fmt, const lzb_compress_options* opt, uint64_t in_len) {
    (void)opt;
    if (fmt == LZB_FMT_LZMA2 || fmt == LZB_FMT_XZ) return lzb_encode_exact(fmt, in_len);
    return in_len + in_len / 4 + 128;
}

static uint32_t host_crc32(const uint8_t* p, size_t n) {
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int b = 0; b < 8; b++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
    }
    return ~c;
}

// caller holds ctx->mu
static int encode_device_locked(lzb_ctx* ctx, int fmt, const lzb_compress_options* opt, const uint8_t* d_in,
                                const uint64_t* in_off, uint32_t n, uint8_t* d_out, const uint64_t* out_off,
                                uint64_t* out_len, lzb_status* st, void* cuda_stream) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    std::vector<LzbEncItem> items(n);
    for (uint32_t i = 0; i < n; i++)
        items[i] = LzbEncItem{in_off[i], in_off[i + 1] - in_off[i], out_off[i], out_off[i + 1] - out_off[i]};
    CUDA_TRY(ctx, ctx->d_enc_items.ensure(n * sizeof(LzbEncItem)));
    CUDA_TRY(ctx, ctx->d_enc_results.ensure(n * sizeof(LzbEncResult)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_enc_items.p, items.data(), n * sizeof(LzbEncItem), cudaMemcpyHostToDevice, s));
    const LzbEncItem* d_items = ctx->d_enc_items.as<LzbEncItem>();
    LzbEncResult* d_res = ctx->d_enc_results.as<LzbEncResult>();
    if (fmt == LZB_FMT_LZMA) {  // K5: one thread per stream, 16 per CTA (12 KB of probabilities each)
        static const lzb_compress_options defaults = {0, 0, {0, 0, 0, 0, 0, 0}, 0};
        const int smem = LZB_ENC_LANES * LZB_ENC_TABLE_U16 * 2;
        if (!ctx->enc_smem_configured) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(lzb_literal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            ctx->enc_smem_configured = 1;
        }
        const uint32_t grid = std::min<uint32_t>((n + LZB_ENC_LANES - 1) / LZB_ENC_LANES, (uint32_t)ctx->sm_count);
        lzb_literal_kernel<<<grid, LZB_ENC_LANES, smem, s>>>(d_items, n, d_in, d_out, opt ? *opt : defaults, d_res);
        CUDA_TRY(ctx, cudaGetLastError());
    } else {  // K4: one CTA per 64 KiB piece of every stream that fits its capacity + one thread per stream for the frame
        std::vector<uint32_t> piece_stream, first_piece(n);
        for (uint32_t i = 0; i < n; i++) {
            first_piece[i] = (uint32_t)piece_stream.size();
            if (lzb_encode_exact(fmt, items[i].in_len) > items[i].out_cap) continue;
            const uint64_t pieces = (items[i].in_len + 0xFFFFull) >> 16;
            piece_stream.insert(piece_stream.end(), (size_t)pieces, i);
        }
        const size_t np = piece_stream.size();
        CUDA_TRY(ctx, ctx->d_enc_pieces.ensure((np + n + 1) * 4));
        uint32_t* d_piece_stream = ctx->d_enc_pieces.as<uint32_t>();
        uint32_t* d_first_piece = d_piece_stream + np;
        if (np) CUDA_TRY(ctx, cudaMemcpyAsync(d_piece_stream, piece_stream.data(), np * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_first_piece, first_piece.data(), n * 4, cudaMemcpyHostToDevice, s));
        LzbXzHead head;  // write_header (xz.rs:31-45) + the block header of write_block (xz.rs:78-97)
        static const uint8_t magic[8] = {0xFD, 0x37, 0x7A, 0x58, 0x5A, 0x00, /* StreamFlags: check None */ 0x00, 0x00};
        static const uint8_t bh[8] = {8 >> 2, 0x00, 0x21, 1, 22, 0, 0, 0};
        memcpy(head.b, magic, 8);
        const uint32_t c1 = host_crc32(magic + 6, 2), c2 = host_crc32(bh, 8);
        for (int i = 0; i < 4; i++) head.b[8 + i] = (uint8_t)(c1 >> (8 * i));
        memcpy(head.b + 12, bh, 8);
        for (int i = 0; i < 4; i++) head.b[20 + i] = (uint8_t)(c2 >> (8 * i));
        if (np) {
            lzb_store_kernel<<<(unsigned)np, 256, 0, s>>>(d_items, d_piece_stream, d_first_piece, d_in, d_out,
                                                          fmt == LZB_FMT_XZ ? 24u : 0u);
            CUDA_TRY(ctx, cudaGetLastError());
        }
        lzb_frame_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_items, n, fmt, d_out, head, d_res);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    std::vector<LzbEncResult> res(n);
    CUDA_TRY(ctx, cudaMemcpyAsync(res.data(), d_res, n * sizeof(LzbEncResult), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    for (uint32_t i = 0; i < n; i++) {
        memset(&st[i], 0, sizeof st[i]);
        st[i].code = res[i].code;
        st[i].kind = res[i].code == LZB_OK ? LZB_KIND_OK : LZB_KIND_INTERNAL;
        st[i].a0 = res[i].code == LZB_OK ? 0 : res[i].out_len;
        out_len[i] = res[i].code == LZB_OK ? res[i].out_len : 0;
    }
    return LZB_RC_OK;
}

extern "C" int lzb_encode_batch_device(lzb_ctx* ctx, int fmt, const lzb_compress_options* opt, const uint8_t* d_in,
                                       const uint64_t* in_off, uint32_t n, uint8_t* d_out, const uint64_t* out_off,
                                       uint64_t* out_len, lzb_status* st, void* cuda_stream) {
    if (!ctx || !in_off || !out_off || !out_len || !st || fmt < 0 || fmt > 2) return LZB_RC_BAD_ARG;
    if (n == 0) return LZB_RC_OK;
    if (!d_in || !d_out) return LZB_RC_BAD_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    return encode_device_locked(ctx, fmt, opt, d_in, in_off, n, d_out, out_off, out_len, st, cuda_stream);
}

extern "C" int lzb_encode_batch(lzb_ctx* ctx, int fmt, const lzb_compress_options* opt, const uint8_t* in,
                                const uint64_t* in_off, uint32_t n, uint8_t* out, const uint64_t* out_off, uint64_t* out_len,
                                lzb_status* st) {
    if (!ctx || !in_off || !out_off || !out_len || !st || fmt < 0 || fmt > 2) return LZB_RC_BAD_ARG;
    if (n == 0) return LZB_RC_OK;
    if (!in || !out) return LZB_RC_BAD_ARG;
    const uint64_t in_lo = in_off[0], in_hi = in_off[n], out_lo = out_off[0], out_hi = out_off[n];
    std::lock_guard<std::mutex> lock(ctx->mu);  // the staging buffers belong to this call from upload to download
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, ctx->d_in.ensure((in_hi - in_lo) + 64));
    CUDA_TRY(ctx, ctx->d_out.ensure((out_hi - out_lo) + 64));
    uint8_t* d_in0 = ctx->d_in.as<uint8_t>() + (in_lo & 15);
    uint8_t* d_out0 = ctx->d_out.as<uint8_t>() + (out_lo & 15);
    if (in_hi > in_lo)
        CUDA_TRY(ctx, cudaMemcpyAsync(d_in0, in + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, ctx->stream));
    int rc = encode_device_locked(ctx, fmt, opt, d_in0 - in_lo, in_off, n, d_out0 - out_lo, out_off, out_len, st, nullptr);
    if (rc != LZB_RC_OK) return rc;
    if (out_hi > out_lo)
        CUDA_TRY(ctx, cudaMemcpyAsync(out + out_lo, d_out0, out_hi - out_lo, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LZB_RC_OK;
}
