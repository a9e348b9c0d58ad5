// emul.cpp -- TEST INFRASTRUCTURE (never linked into the product library).
// Builds K1's decode core (lzma_rs_b200/csrc/lzb_decode_core.h) as plain C++ with a 1-lane "warp", plugs it
// into the product's host planning code (lzb_plan.cpp) through the Executor interface, and exposes the same
// batch entry point as the C ABI.  This lets the CPU-only test tier check the decode logic, the container walk
// and the status mapping against the oracle without a GPU.  The real GPU parity tests are tests/test_gpu_*.py.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../lzma_rs_b200/csrc/lzb_decode_core.h"
#include "../../lzma_rs_b200/csrc/lzb_plan.h"

namespace {
class HostEmulExecutor : public lzb::Executor {
   public:
    HostEmulExecutor(const uint8_t* in, uint8_t* out) : in_(in), out_(out) {}
    int decode(const LzbItem* items, uint32_t n, uint32_t max_lclp, uint64_t /*stored_bytes*/, LzbResult* results) override {
        const uint32_t small_lclp = max_lclp > 4 ? 4 : max_lclp;
        std::vector<uint16_t> T(lzb_table_u16(small_lclp) + 8), M(lzb_matched_u16(small_lclp) + 8), T4, M4, G, TL;
        for (uint32_t i = 0; i < n; i++) {
            memset(&results[i], 0, sizeof results[i]);
            const uint32_t lclp = (uint32_t)items[i].lc + items[i].lp;
            if (items[i].kind == LZB_ITEM_LZMA && lclp > 4) {  // whole literal table outside "shared memory"
                G.assign((size_t)0x300u << lclp, 0);
                run<true>(items + i, T.data(), G.data(), lclp, results + i);
                continue;
            }
            if (getenv("LZB_EMUL_LAT")) {  // the latency form of K1 (look-ahead walks, whole literal table in "shared memory")
                TL.assign(lzb_lat_table_u16(small_lclp) + 8, 0);
                run_lat(items + i, TL.data(), small_lclp, results + i);
                if (results[i].code == LZB_E_UNSUPPORTED && results[i].a1 == small_lclp && results[i].a0 <= 4) {
                    TL.assign(lzb_lat_table_u16(4) + 8, 0);
                    run_lat(items + i, TL.data(), 4, results + i);
                }
                continue;
            }
            run<false>(items + i, T.data(), M.data(), small_lclp, results + i);
            if (results[i].code == LZB_E_UNSUPPORTED && results[i].a1 == small_lclp && results[i].a0 <= 4) {
                T4.resize(lzb_table_u16(4) + 8);  // framing scan under-estimated lc+lp: retry with the LZMA2 maximum
                M4.resize(lzb_matched_u16(4) + 8);
                run<false>(items + i, T4.data(), M4.data(), 4, results + i);
            }
        }
        return LZB_RC_OK;
    }
    int crc(const lzb::CrcRange* r, uint32_t n, uint32_t* c32, uint64_t* c64) override {
        static uint32_t t32[256];
        static uint64_t t64[256];
        static bool init = false;
        if (!init) {
            for (uint32_t i = 0; i < 256; i++) {
                uint32_t c = i;
                uint64_t e = i;
                for (int k = 0; k < 8; k++) {
                    c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
                    e = (e & 1) ? (e >> 1) ^ 0xC96C5795D7870F42ull : e >> 1;
                }
                t32[i] = c;
                t64[i] = e;
            }
            init = true;
        }
        for (uint32_t i = 0; i < n; i++) {
            uint32_t c = 0xFFFFFFFFu;
            uint64_t e = ~0ull;
            for (uint64_t k = 0; k < r[i].len; k++) {
                uint8_t b = out_[r[i].off + k];
                c = t32[(c ^ b) & 0xFF] ^ (c >> 8);
                e = t64[(e ^ b) & 0xFF] ^ (e >> 8);
            }
            c32[i] = c ^ 0xFFFFFFFFu;
            c64[i] = ~e;
        }
        return LZB_RC_OK;
    }

    int scratch(uint64_t bytes, uint64_t* off) override {
        scratch_.emplace_back(bytes + 64);
        *off = (uint64_t)(scratch_.back().data() - out_);  // output-blob coordinates
        return LZB_RC_OK;
    }
    int read_out(uint64_t off, uint64_t len, uint8_t* dst) override {
        memcpy(dst, out_ + off, len);
        return LZB_RC_OK;
    }

   private:
    template <bool BIG>
    void run(const LzbItem* it, uint16_t* T, uint16_t* G, uint32_t lclp, LzbResult* res) {
        const LzbKC kc = LZB_KC_INIT;
        const TabPtr tab = {T};
        const TabPtr plain = {BIG ? G : T + T_LIT};
        const TabPtr matched = {BIG ? G + 0x100 : G};
        decode_item<BIG, false, 1>(it, in_, out_, T, G, tab, plain, matched, kc, lclp, res, 0);
    }
    void run_lat(const LzbItem* it, uint16_t* T, uint32_t lclp, LzbResult* res) {
        const LzbKC kc = LZB_KC_INIT;
        const TabPtr tab = {T};
        const TabPtr plain = {T + T_LIT};
        const TabPtr matched = {T + T_LIT + 0x100};
        decode_item<false, false, 0, false, true>(it, in_, out_, T, nullptr, tab, plain, matched, kc, lclp, res, 0);
    }
    const uint8_t* in_;
    uint8_t* out_;
    std::vector<std::vector<uint8_t>> scratch_;
};
}  // namespace

extern "C" int emul_decode_batch(int fmt, const lzb_options* opt, const uint8_t* in, const uint64_t* in_off, uint32_t n,
                                 uint8_t* out, const uint64_t* out_off, uint64_t* out_len, uint64_t* consumed,
                                 lzb_status* st) {
    HostEmulExecutor ex(in, out);
    std::vector<lzb::StreamOut> outs(n);
    int rc = lzb::decode_batch(ex, fmt, opt, in, in_off, n, out_off, outs.data());
    for (uint32_t i = 0; i < n; i++) {
        out_len[i] = outs[i].out_len;
        consumed[i] = outs[i].consumed;
        st[i] = outs[i].st;
    }
    return rc;
}

extern "C" uint64_t emul_scan_capacity(int fmt, const lzb_options* opt, const uint8_t* p, uint64_t len) {
    return lzb::scan_capacity(fmt, opt, p, len);
}

// ---- placement planner (lzb_sched.h, host-only C++): exposed so the CPU tier can check it against its Python twin ----
#include "../../lzma_rs_b200/csrc/lzb_sched.h"

// work[n] in queue order (longest first).  Writes the order array (capacity n + sms * warps) and returns its length;
// info = {throttled, n_static, grid, parked}; model = {plain makespan, predicted makespan}.
extern "C" uint32_t emul_sched_plan(const double* work, uint32_t n, uint32_t sms, uint32_t warps, uint32_t* order_out,
                                    uint32_t* info, double* model) {
    std::vector<uint32_t> sorted(n);
    std::vector<double> w(work, work + n);
    for (uint32_t i = 0; i < n; i++) sorted[i] = i;
    lzb_sched::Plan p = lzb_sched::plan(sorted, w, sms, warps);
    for (size_t k = 0; k < p.order.size(); k++) order_out[k] = p.order[k];
    info[0] = p.throttled;
    info[1] = p.n_static;
    info[2] = p.grid;
    info[3] = p.parked;
    model[0] = p.plain;
    model[1] = p.predicted;
    return (uint32_t)p.order.size();
}

extern "C" double emul_sched_simulate(const double* work, uint32_t n, const uint32_t* counts, uint32_t n_counts, uint32_t sms,
                                      uint32_t warps) {
    std::vector<double> w(work, work + n);
    if (!counts) return lzb_sched::simulate(w, nullptr, sms, warps);
    std::vector<uint32_t> c(counts, counts + n_counts);
    return lzb_sched::simulate(w, &c, sms, warps);
}

// ---- decompress::raw decoder objects: K1's CARRY path on the CPU (the state record lives in host memory here) ----
// Same contract as lzb_raw_create / reset / decompress of the C ABI, minus the device: one record per decoder, every
// call runs decode_item<LIT_GLOBAL, ..., CARRY> on a copy and commits it unless the output capacity was too small.
struct EmulRaw {
    int fmt;
    uint32_t lc, lp, pb, dict_size, lclp_cap;
    std::vector<uint8_t> state;
};
static void emul_raw_fresh(EmulRaw* r) {
    LzbCarry h;
    memset(&h, 0, sizeof h);
    h.fresh = 1;
    h.lc = r->fmt == LZB_FMT_LZMA ? r->lc : 0;
    h.lp = r->fmt == LZB_FMT_LZMA ? r->lp : 0;
    h.pb = r->fmt == LZB_FMT_LZMA ? r->pb : 0;
    h.lclp_cap = r->lclp_cap;
    memcpy(r->state.data(), &h, sizeof h);
}
extern "C" EmulRaw* emul_raw_create(int fmt, uint32_t lc, uint32_t lp, uint32_t pb, uint32_t dict_size) {
    EmulRaw* r = new EmulRaw{fmt, lc, lp, pb, dict_size, fmt == LZB_FMT_LZMA ? lc + lp : 4u, {}};
    r->state.assign((size_t)lzb_carry_bytes(r->lclp_cap) + 64, 0);
    emul_raw_fresh(r);
    return r;
}
extern "C" void emul_raw_reset(EmulRaw* r) { emul_raw_fresh(r); }
extern "C" void emul_raw_destroy(EmulRaw* r) { delete r; }
// out: caller's buffer of `cap` bytes.  Returns the status code (LZB_E_CAPACITY: nothing committed, call again larger).
extern "C" int emul_raw_decompress(EmulRaw* r, const lzb_options* opt, const uint8_t* in, uint64_t in_len, uint8_t* out,
                                   uint64_t cap, uint64_t* out_len, uint64_t* consumed, lzb_status* st) {
    std::vector<uint8_t> work(r->state);
    std::vector<uint16_t> T(T_LIT + 16);
    std::vector<uint8_t> inbuf(in, in + in_len);
    inbuf.resize(in_len + 16);
    LzbItem it;
    memset(&it, 0, sizeof it);
    it.in_len = in_len;
    it.out_cap = cap;
    it.unpacked = (r->fmt == LZB_FMT_LZMA && opt && opt->has_provided) ? opt->provided : LZB_UNKNOWN_SIZE;
    it.memlimit = opt && opt->has_memlimit ? opt->memlimit : ~0ull;
    it.dict_size = r->dict_size;
    it.kind = r->fmt == LZB_FMT_LZMA ? LZB_ITEM_LZMA : LZB_ITEM_LZMA2;
    it.lc = (uint8_t)r->lc;
    it.lp = (uint8_t)r->lp;
    it.pb = (uint8_t)r->pb;
    it.flags = LZB_ITEM_F_CARRY;
    it.host_out = (uint64_t)(uintptr_t)work.data();
    LzbResult res;
    memset(&res, 0, sizeof res);
    const LzbKC kc = LZB_KC_INIT;
    uint16_t* lit = reinterpret_cast<uint16_t*>(work.data() + sizeof(LzbCarry)) + T_LIT;
    const TabPtr tab = {T.data()}, plain = {lit}, matched = {lit + 0x100};
    decode_item<true, false, 0, true>(&it, inbuf.data(), out, T.data(), lit, tab, plain, matched, kc, r->lclp_cap, &res, 0);
    if (res.code == LZB_OK) r->state.swap(work);  // like lzb_raw_decompress: a failed call commits nothing
    lzb::status_from_result(res, st);
    *out_len = res.sink_len;
    *consumed = res.consumed;
    return res.code;
}
