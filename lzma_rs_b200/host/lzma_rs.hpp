// lzma_rs.hpp -- C++ host-side mirror of the reference's decode API over the C ABI (include/lzma_b200.h).
//
//   lzma_rs::lzma_decompress(std::istream&, std::ostream&)                     src/lib.rs:44-49
//   lzma_rs::lzma_decompress_with_options(in, out, decompress::Options)        src/lib.rs:52-60
//   lzma_rs::lzma2_decompress(in, out)                                          src/lib.rs:83-88
//   lzma_rs::xz_decompress(in, out)                                             src/lib.rs:100-105
//
// Errors are thrown as lzma_rs::error::Error whose what() is the reference's Display string and whose `kind`
// is the error::Error variant.  Partial output is written before the throw, as in the reference.
// Header-only; link with -llzma_b200.  There is no CPU fallback.
#pragma once
#include <algorithm>
#include <cstdint>
#include <istream>
#include <iterator>
#include <optional>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "lzma_b200.h"

namespace lzma_rs {

namespace error {
enum class Kind { IoError = 1, HeaderTooShort = 2, LzmaError = 3, XzError = 4, Internal = 5 };
struct Error : std::runtime_error {
    Kind kind;
    lzb_status status;
    Error(Kind k, const std::string& display, const lzb_status& st) : std::runtime_error(display), kind(k), status(st) {}
};
}  // namespace error

namespace decompress {
struct UnpackedSize {  // options.rs:24-43
    enum Mode { ReadFromHeader = 0, ReadHeaderButUseProvided = 1, UseProvided = 2 } mode = ReadFromHeader;
    std::optional<uint64_t> value;
};
struct Options {  // options.rs:3-20
    UnpackedSize unpacked_size;
    std::optional<size_t> memlimit;
    bool allow_incomplete = false;  // stream API only
};
}  // namespace decompress

namespace detail {
inline lzb_ctx* ctx() {
    static lzb_ctx* c = [] {
        lzb_ctx* p = nullptr;
        if (lzb_create(&p, -1) != LZB_RC_OK) throw std::runtime_error("lzb_create failed: no CUDA device (no CPU fallback)");
        return p;
    }();
    return c;
}
inline void run(int fmt, const lzb_options* opt, std::istream& in, std::ostream& out) {
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    uint8_t* o = nullptr;
    size_t out_len = 0, consumed = 0;
    lzb_status st{};
    int rc = lzb_decompress_alloc(ctx(), fmt, opt, buf.data(), buf.size(), &o, &out_len, &consumed, &st);
    if (rc != LZB_RC_OK) throw std::runtime_error(std::string("lzma_b200: ") + lzb_last_error(ctx()));
    if (out_len) out.write(reinterpret_cast<const char*>(o), (std::streamsize)out_len);
    lzb_free(o);
    if (st.code != LZB_OK) {
        char msg[512];
        lzb_format_error(&st, msg, sizeof msg);
        throw error::Error(static_cast<error::Kind>(st.kind), msg, st);
    }
    in.clear();
    in.seekg((std::streamoff)consumed - (std::streamoff)buf.size(), std::ios::cur);  // unread trailing bytes stay
    out.flush();
}
}  // namespace detail

namespace detail {
inline void run_lzma(std::istream& in, std::ostream& out, const decompress::Options& o, bool allow_incomplete) {
    lzb_options n{};
    n.unpacked_mode = (uint8_t)o.unpacked_size.mode;
    n.has_provided = o.unpacked_size.value.has_value();
    n.provided = o.unpacked_size.value.value_or(0);
    n.has_memlimit = o.memlimit.has_value();
    n.memlimit = o.memlimit.value_or(0);
    n.allow_incomplete = allow_incomplete;
    run(LZB_FMT_LZMA, &n, in, out);
}
}  // namespace detail
// allow_incomplete is an option of the stream API only (options.rs:15-19): the one-shot decoder ignores it
inline void lzma_decompress_with_options(std::istream& in, std::ostream& out, const decompress::Options& o) {
    detail::run_lzma(in, out, o, false);
}
inline void lzma_decompress(std::istream& in, std::ostream& out) { lzma_decompress_with_options(in, out, {}); }

// lzma_rs::{lzma_compress, lzma_compress_with_options, lzma2_compress, xz_compress} (src/lib.rs:63-80, 91-97, 108-110).
// The reference's encoders are format writers (literals only / stored chunks only); the GPU writes the same bytes.
namespace compress {
struct UnpackedSize {  // src/encode/options.rs:10-24
    bool skip_writing_to_header = false;
    std::optional<uint64_t> value;  // WriteToHeader(value)
};
struct Options {
    UnpackedSize unpacked_size;
};
}  // namespace compress
namespace detail {
inline void encode(int fmt, const lzb_compress_options* opt, std::istream& in, std::ostream& out) {
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    uint64_t cap = lzb_encode_bound(fmt, opt, buf.size());
    for (;;) {
        std::vector<uint8_t> o((size_t)cap + 16);
        const uint64_t in_off[2] = {0, buf.size()}, out_off[2] = {0, cap};
        uint64_t out_len = 0;
        lzb_status st{};
        static const uint8_t none = 0;
        int rc = lzb_encode_batch(ctx(), fmt, opt, buf.empty() ? &none : buf.data(), in_off, 1, o.data(), out_off, &out_len, &st);
        if (rc != LZB_RC_OK) throw std::runtime_error(std::string("lzma_b200: ") + lzb_last_error(ctx()));
        if (st.code == LZB_E_CAPACITY) {
            cap = st.a0;
            continue;
        }
        out.write(reinterpret_cast<const char*>(o.data()), (std::streamsize)out_len);
        out.flush();
        return;
    }
}
}  // namespace detail
inline void lzma_compress_with_options(std::istream& in, std::ostream& out, const compress::Options& o) {
    lzb_compress_options n{};
    n.skip_size_field = o.unpacked_size.skip_writing_to_header;
    n.has_value = !o.unpacked_size.skip_writing_to_header && o.unpacked_size.value.has_value();
    n.value = o.unpacked_size.value.value_or(0);
    detail::encode(LZB_FMT_LZMA, &n, in, out);
}
inline void lzma_compress(std::istream& in, std::ostream& out) { lzma_compress_with_options(in, out, {}); }
inline void lzma2_compress(std::istream& in, std::ostream& out) { detail::encode(LZB_FMT_LZMA2, nullptr, in, out); }
inline void xz_compress(std::istream& in, std::ostream& out) { detail::encode(LZB_FMT_XZ, nullptr, in, out); }

// lzma_rs::decompress::raw (feature `raw_decoder`, src/lib.rs:29-35): decoder objects over lzb_raw_*.
namespace decompress {
namespace raw {
struct LzmaProperties {  // lzma.rs:41-66
    uint32_t lc, lp, pb;
    void validate() const {
        if (lc > 8 || lp > 4 || pb > 4) throw std::invalid_argument("LzmaProperties: lc <= 8, lp <= 4, pb <= 4");
    }
};
struct LzmaParams {  // LzmaParams::new, lzma.rs:68-93
    LzmaProperties properties;
    uint32_t dict_size;
    std::optional<uint64_t> unpacked_size;
};
namespace detail {
// One decode through a raw decoder object (lzb_raw_decompress): output, error and "unread trailing bytes stay" like run().
inline void run_raw(lzb_raw* r, const lzb_options* opt, std::istream& in, std::ostream& out) {
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    uint8_t* o = nullptr;
    size_t out_len = 0, consumed = 0;
    lzb_status st{};
    int rc = lzb_raw_decompress(r, opt, buf.data(), buf.size(), &o, &out_len, &consumed, &st);
    if (rc != LZB_RC_OK) throw std::runtime_error(std::string("lzma_b200: ") + lzb_last_error(lzma_rs::detail::ctx()));
    if (out_len) out.write(reinterpret_cast<const char*>(o), (std::streamsize)out_len);
    lzb_free(o);
    if (st.code != LZB_OK) {
        char msg[512];
        lzb_format_error(&st, msg, sizeof msg);
        throw error::Error(static_cast<error::Kind>(st.kind), msg, st);
    }
    in.clear();
    in.seekg((std::streamoff)consumed - (std::streamoff)buf.size(), std::ios::cur);
    out.flush();
}
}  // namespace detail
// Like the reference's, both decoders keep their DecoderState (probabilities, state, rep distances; LZMA2: the properties
// of the last props reset) from one decompress() to the next until reset(); the state lives on the device (lzb_raw_*).
class LzmaDecoder {  // lzma.rs:597-648: `in` is a headerless LZMA stream
   public:
    LzmaDecoder(const LzmaParams& p, std::optional<size_t> memlimit) : p_(p), memlimit_(memlimit), size_(p.unpacked_size) {
        p.properties.validate();
        if (lzb_raw_create(lzma_rs::detail::ctx(), LZB_FMT_LZMA, p.properties.lc, p.properties.lp, p.properties.pb, p.dict_size,
                           &raw_) != LZB_RC_OK)
            throw std::invalid_argument("raw LzmaDecoder: lzb_raw_create failed");
    }
    ~LzmaDecoder() { lzb_raw_destroy(raw_); }
    LzmaDecoder(const LzmaDecoder&) = delete;
    LzmaDecoder& operator=(const LzmaDecoder&) = delete;
    void reset() { lzb_raw_reset(raw_); }  // reset(None): the size stays (lzma.rs:620-627)
    void reset(std::optional<uint64_t> unpacked_size) {
        size_ = unpacked_size;
        lzb_raw_reset(raw_);
    }
    void decompress(std::istream& in, std::ostream& out) {
        lzb_options n{};
        n.unpacked_mode = 2;
        n.has_provided = size_.has_value();
        n.provided = size_.value_or(0);
        n.has_memlimit = memlimit_.has_value();
        n.memlimit = memlimit_.value_or(0);
        detail::run_raw(raw_, &n, in, out);
    }

   private:
    LzmaParams p_;
    std::optional<size_t> memlimit_;
    std::optional<uint64_t> size_;
    lzb_raw* raw_ = nullptr;
};
class Lzma2Decoder {  // lzma2.rs:11-82
   public:
    Lzma2Decoder() {
        if (lzb_raw_create(lzma_rs::detail::ctx(), LZB_FMT_LZMA2, 0, 0, 0, 0, &raw_) != LZB_RC_OK)
            throw std::runtime_error("raw Lzma2Decoder: lzb_raw_create failed");
    }
    ~Lzma2Decoder() { lzb_raw_destroy(raw_); }
    Lzma2Decoder(const Lzma2Decoder&) = delete;
    Lzma2Decoder& operator=(const Lzma2Decoder&) = delete;
    void reset() { lzb_raw_reset(raw_); }
    void decompress(std::istream& in, std::ostream& out) { detail::run_raw(raw_, nullptr, in, out); }

   private:
    lzb_raw* raw_ = nullptr;
};
}  // namespace raw
}  // namespace decompress
inline void lzma2_decompress(std::istream& in, std::ostream& out) { detail::run(LZB_FMT_LZMA2, nullptr, in, out); }

// lzma_rs::decompress::Stream (feature `stream`, src/decode/stream.rs:66-346) as a buffering facade: write() stores the
// bytes (and rejects an invalid properties byte at once, like the reference), finish() decodes the whole stream on the
// GPU and hands everything to the sink.  Same results as the reference's incremental decoder on valid input; data
// errors surface in finish().  With allow_incomplete on an unknown-size stream it returns every byte of every complete
// symbol, a few bytes more than the reference (lzma_rs_b200.Stream._incomplete_output shows how to trim).
namespace decompress {
class Stream {
   public:
    explicit Stream(std::ostream& out, const Options& o = {}) : out_(&out), opt_(o) {}
    size_t write(const void* data, size_t n) {
        if (!out_) return 0;  // a previous write failed (stream.rs:230,310)
        const bool first = buf_.empty();
        buf_.append(static_cast<const char*>(data), n);
        if (first && !buf_.empty() && (unsigned char)buf_[0] >= 225) {
            const unsigned props = (unsigned char)buf_[0];
            out_ = nullptr;
            lzb_status st{};
            st.code = LZB_E_LZMA_PROPS;
            st.kind = LZB_KIND_LZMA;
            st.a0 = props;
            throw error::Error(error::Kind::LzmaError,
                               "lzma error: LZMA header invalid properties: " + std::to_string(props) + " must be < 225", st);
        }
        return n;
    }
    std::ostream& finish() {
        lzb_status st{};
        if (!out_) throw error::Error(error::Kind::LzmaError, "lzma error: can't finish stream because of previous write error", st);
        std::ostream& out = *out_;
        out_ = nullptr;
        if (buf_.empty()) return out;
        const size_t hdr = opt_.unpacked_size.mode == UnpackedSize::UseProvided ? 5 : 13;
        if (buf_.size() < hdr + 5) throw error::Error(error::Kind::LzmaError, "lzma error: failed to read header", st);
        std::istringstream in(buf_);
        detail::run_lzma(in, out, opt_, opt_.allow_incomplete);
        return out;
    }

   private:
    std::ostream* out_;
    Options opt_;
    std::string buf_;
};
}  // namespace decompress
inline void xz_decompress(std::istream& in, std::ostream& out) { detail::run(LZB_FMT_XZ, nullptr, in, out); }

// ---- batch forms (n independent streams per call: what the benchmark drives) and several GPUs ----
namespace batch {
struct Input {
    const uint8_t* data;
    size_t size;
};
struct Result {
    std::vector<uint8_t> data;  // what the reference would have written to its sink (also on error)
    size_t consumed = 0;
    lzb_status status{};
    std::string error;  // "" or the reference's Display string
    bool ok() const { return status.code == LZB_OK; }
};
namespace detail {
inline lzb_multi*& multi() {
    static lzb_multi* m = nullptr;
    return m;
}
inline std::vector<Result> run(int fmt, const lzb_options* opt, const std::vector<Input>& in) {
    const uint32_t n = (uint32_t)in.size();
    std::vector<Result> res(n);
    if (!n) return res;
    std::vector<uint64_t> in_off(n + 1, 0), out_off(n + 1, 0), cap(n), out_len(n), consumed(n);
    std::vector<uint8_t> blob;
    for (uint32_t i = 0; i < n; i++) {
        blob.insert(blob.end(), in[i].data, in[i].data + in[i].size);
        in_off[i + 1] = blob.size();
    }
    blob.resize(blob.size() + 16);
    lzb_ctx* c = multi() ? lzb_multi_ctx(multi(), 0) : lzma_rs::detail::ctx();
    if (lzb_scan(c, fmt, opt, blob.data(), in_off.data(), n, cap.data()) != LZB_RC_OK) throw std::runtime_error("lzb_scan failed");
    std::vector<uint32_t> pending(n);
    for (uint32_t i = 0; i < n; i++) pending[i] = i;
    while (!pending.empty()) {  // streams of unknown size that report LZB_E_CAPACITY are decoded again, larger
        const uint32_t m = (uint32_t)pending.size();
        std::vector<uint64_t> sin(m + 1, 0), sout(m + 1, 0), slen(m), scons(m);
        std::vector<uint8_t> sblob;
        for (uint32_t k = 0; k < m; k++) {
            sblob.insert(sblob.end(), in[pending[k]].data, in[pending[k]].data + in[pending[k]].size);
            sin[k + 1] = sblob.size();
            sout[k + 1] = sout[k] + ((cap[pending[k]] + 15) & ~15ull);
        }
        sblob.resize(sblob.size() + 16);
        std::vector<uint8_t> out(sout[m] + 16);
        std::vector<lzb_status> st(m);
        const int rc = multi() ? lzb_decode_batch_multi(multi(), fmt, opt, sblob.data(), sin.data(), m, out.data(), sout.data(),
                                                        slen.data(), scons.data(), st.data(), nullptr)
                               : lzb_decode_batch(c, fmt, opt, sblob.data(), sin.data(), m, out.data(), sout.data(), slen.data(),
                                                  scons.data(), st.data());
        if (rc != LZB_RC_OK) throw std::runtime_error(std::string("lzma_b200: ") + lzb_last_error(c));
        std::vector<uint32_t> again;
        for (uint32_t k = 0; k < m; k++) {
            const uint32_t i = pending[k];
            if (st[k].code == LZB_E_CAPACITY && cap[i] < 0xFFFFF000ull) {
                cap[i] = std::min<uint64_t>(std::max<uint64_t>(cap[i] * 2, st[k].a0 + 65536), 0xFFFFF000ull);
                again.push_back(i);
                continue;
            }
            res[i].data.assign(out.begin() + sout[k], out.begin() + sout[k] + slen[k]);
            res[i].consumed = scons[k];
            res[i].status = st[k];
            if (st[k].code != LZB_OK) {
                char msg[512];
                lzb_format_error(&st[k], msg, sizeof msg);
                res[i].error = msg;
            }
        }
        pending.swap(again);
    }
    return res;
}
}  // namespace detail
// Spread the batch calls over these CUDA devices (empty = every visible device): lzb_create_multi.
inline void set_devices(const std::vector<int>& devices) {
    if (detail::multi()) lzb_destroy_multi(detail::multi());
    if (lzb_create_multi(&detail::multi(), devices.empty() ? nullptr : devices.data(), (int)devices.size()) != LZB_RC_OK)
        throw std::runtime_error("lzb_create_multi failed: no CUDA device (no CPU fallback)");
}
inline std::vector<Result> lzma_decompress_batch(const std::vector<Input>& in) { return detail::run(LZB_FMT_LZMA, nullptr, in); }
inline std::vector<Result> lzma2_decompress_batch(const std::vector<Input>& in) { return detail::run(LZB_FMT_LZMA2, nullptr, in); }
inline std::vector<Result> xz_decompress_batch(const std::vector<Input>& in) { return detail::run(LZB_FMT_XZ, nullptr, in); }
}  // namespace batch

}  // namespace lzma_rs
