// lzb_encode_kernels.cu -- the compress side of the reference API on the GPU.
//
// gendx/lzma-rs's encoders are format writers, not compressors (README.md:29-32 of the reference):
//   lzma2_compress / xz_compress  emit stored chunks only      (src/encode/lzma2.rs:4-26, src/encode/xz.rs:9-183)
//   lzma_compress                 emits literals only           (src/encode/dumbencoder.rs:24-139)
// so that is what these kernels write, byte for byte:
//   K4 lzb_store_kernel  : one CTA per 64 KiB piece -- 3-byte chunk header + a misaligned 64 KiB copy (HBM-bound)
//   K4 lzb_frame_kernel  : one thread per stream -- LZMA2 terminator; XZ stream header, block header, padding, index, footer
//   K5 lzb_literal_kernel: one THREAD per stream.  Unlike decoding, every stream runs the same control flow (one
//                          is_match decision and an 8-level literal walk per input byte), so streams map to lanes
//                          without divergence; each lane owns a 12 KB probability table in shared memory (16 per SM).
#include <cuda_runtime.h>
#include <stdint.h>

#include "lzb_decode_core.h"  // warp_copy
#include "lzb_encode.h"
#include "lzb_types.h"

// ------------------------------------------------------------------------------------------------
// K4: stored chunks (lzma2.rs:4-26: status 1, BE u16 size - 1, payload; 0x00 after the last chunk)
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256)
    lzb_store_kernel(const LzbEncItem* __restrict__ items, const uint32_t* __restrict__ piece_stream,
                     const uint32_t* __restrict__ first_piece, const uint8_t* __restrict__ in, uint8_t* out,
                     uint32_t prefix) {
    const uint32_t piece = blockIdx.x;
    const uint32_t s = piece_stream[piece];
    const uint32_t j = piece - first_piece[s];
    const LzbEncItem it = items[s];
    const uint64_t lo = (uint64_t)j << 16;
    const uint32_t k = (uint32_t)(it.in_len - lo < 0x10000ull ? it.in_len - lo : 0x10000ull);
    const uint8_t* src = in + it.in_off + lo;
    uint8_t* dst = out + it.out_off + prefix + (uint64_t)j * (0x10000ull + 3);
    if (threadIdx.x == 0) {
        dst[0] = 1;  // "uncompressed, reset dict" for EVERY chunk, as the reference writes it
        dst[1] = (uint8_t)((k - 1) >> 8);
        dst[2] = (uint8_t)(k - 1);
    }
    const int lane = threadIdx.x & 31;
    const uint32_t w0 = (threadIdx.x >> 5) * 8192u;  // 8 warps x 8 KiB
    if (w0 < k) warp_copy<true>(dst + 3 + w0, src + w0, k - w0 < 8192u ? k - w0 : 8192u, lane);
}

__device__ uint32_t crc32_bytes(const uint8_t* p, uint32_t n) {  // CRC-32/ISO-HDLC, a few bytes: bitwise
    uint32_t c = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int b = 0; b < 8; b++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
    }
    return ~c;
}

__device__ uint32_t put_multibyte(uint8_t* p, uint64_t v) {  // xz.rs:166-183
    uint32_t n = 0;
    for (;;) {
        const uint8_t b = (uint8_t)(v & 0x7F);
        v >>= 7;
        if (v == 0) {
            p[n++] = b;
            return n;
        }
        p[n++] = 0x80 | b;
    }
}

// fmt LZB_FMT_LZMA2: writes the 0x00 terminator.  fmt LZB_FMT_XZ (encode_stream, xz.rs:9-29): the 24 bytes before
// the chunks (stream header + block header, constant: `head`), then terminator, block padding, index and footer.
extern "C" __global__ void lzb_frame_kernel(const LzbEncItem* __restrict__ items, uint32_t n, int fmt, uint8_t* out,
                                            const LzbXzHead head, LzbEncResult* results) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const LzbEncItem it = items[s];
    LzbEncResult r;
    r.code = LZB_OK;
    r.pad = 0;
    r.out_len = lzb_encode_exact(fmt, it.in_len);
    if (r.out_len > it.out_cap) {  // the host launched no pieces for this stream either
        r.code = LZB_E_CAPACITY;
        results[s] = r;
        return;
    }
    const uint64_t pieces = (it.in_len + 0xFFFFull) >> 16;
    const uint64_t l2 = it.in_len + 3 * pieces + 1;  // LZMA2 bytes incl. the terminator
    uint8_t* o = out + it.out_off;
    if (fmt == LZB_FMT_LZMA2) {
        o[l2 - 1] = 0;
    } else {
        for (int i = 0; i < 24; i++) o[i] = head.b[i];
        uint8_t* p = o + 24 + l2;
        p[-1] = 0;
        const uint64_t unpadded = 12 + l2;  // block header + its CRC + LZMA2 stream (write_block, xz.rs:70-122)
        for (uint64_t pad = ((unpadded ^ 3) + 1) & 3; pad; pad--) *p++ = 0;
        uint8_t idx[32];  // write_index, xz.rs:124-164
        uint32_t k = 0;
        idx[k++] = 0;
        k += put_multibyte(idx + k, 1);
        k += put_multibyte(idx + k, unpadded);
        k += put_multibyte(idx + k, it.in_len);
        for (uint32_t pad = ((k ^ 3) + 1) & 3; pad; pad--) idx[k++] = 0;
        const uint32_t icrc = crc32_bytes(idx, k);
        for (int i = 0; i < 4; i++) idx[k++] = (uint8_t)(icrc >> (8 * i));
        for (uint32_t i = 0; i < k; i++) *p++ = idx[i];
        uint8_t fb[6];  // write_footer, xz.rs:47-68
        const uint32_t backward = (k >> 2) - 1;
        for (int i = 0; i < 4; i++) fb[i] = (uint8_t)(backward >> (8 * i));
        fb[4] = head.b[6];
        fb[5] = head.b[7];
        const uint32_t fcrc = crc32_bytes(fb, 6);
        for (int i = 0; i < 4; i++) *p++ = (uint8_t)(fcrc >> (8 * i));
        for (int i = 0; i < 6; i++) *p++ = fb[i];
        *p++ = 0x59;
        *p++ = 0x5A;
    }
    results[s] = r;
}

// ------------------------------------------------------------------------------------------------
// K5: literal-only LZMA encoder (dumbencoder.rs:24-139 over the range encoder of encode/rangecoder.rs:7-106)
// ------------------------------------------------------------------------------------------------
struct RangeEnc {
    uint32_t range;
    uint64_t low;
    uint32_t cache, cachesz;
    uint8_t* out;       // stream output base
    uint64_t pos, cap;  // bytes produced so far (stores stop at cap, counting goes on: the caller learns the size needed)
};

__device__ __forceinline__ void re_put(RangeEnc& e, uint8_t b) {
    if (e.pos < e.cap) e.out[e.pos] = b;
    e.pos++;
}

__device__ __forceinline__ void re_write_low(RangeEnc& e) {  // rangecoder.rs:34-52
    if (e.low < 0xFF000000ull || e.low > 0xFFFFFFFFull) {
        uint32_t tmp = e.cache;
        for (;;) {
            re_put(e, (uint8_t)(tmp + (uint32_t)(e.low >> 32)));
            tmp = 0xFF;
            if (--e.cachesz == 0) break;
        }
        e.cache = (uint32_t)(e.low >> 24) & 0xFF;
    }
    e.cachesz++;
    e.low = (e.low << 8) & 0xFFFFFFFFull;
}

__device__ __forceinline__ void re_encode_bit(RangeEnc& e, uint16_t* prob, uint32_t bit) {  // rangecoder.rs:63-106
    const uint32_t p = *prob;
    const uint32_t bound = (e.range >> 11) * p;
    if (bit) {
        *prob = (uint16_t)(p - (p >> 5));
        e.low += bound;
        e.range -= bound;
    } else {
        *prob = (uint16_t)(p + ((0x800u - p) >> 5));
        e.range = bound;
    }
    while (e.range < 0x01000000u) {
        e.range <<= 8;
        re_write_low(e);
    }
}

extern "C" __global__ void __launch_bounds__(LZB_ENC_LANES)
    lzb_literal_kernel(const LzbEncItem* __restrict__ items, uint32_t n, const uint8_t* __restrict__ in, uint8_t* out,
                       const lzb_compress_options opt, LzbEncResult* results) {
    extern __shared__ __align__(16) uint16_t enc_smem[];
    uint16_t* T = enc_smem + (size_t)threadIdx.x * LZB_ENC_TABLE_U16;  // [8][0x300] literal probs, then is_match[4]
    for (uint32_t s = blockIdx.x * LZB_ENC_LANES + threadIdx.x; s < n; s += gridDim.x * LZB_ENC_LANES) {
        const LzbEncItem it = items[s];
        for (uint32_t i = 0; i < LZB_ENC_TABLE_U16; i++) T[i] = 0x400;
        uint16_t* is_match = T + 8 * 0x300;
        RangeEnc e;
        e.range = 0xFFFFFFFFu;
        e.low = 0;
        e.cache = 0;
        e.cachesz = 1;
        e.out = out + it.out_off;
        e.pos = 0;
        e.cap = it.out_cap;
        // Encoder::from_stream, dumbencoder.rs:24-62
        re_put(e, (uint8_t)(3 + 9 * (0 + 5 * 2)));
        for (int i = 0; i < 4; i++) re_put(e, (uint8_t)(0x00800000u >> (8 * i)));
        if (!opt.skip_size_field) {
            const uint64_t v = opt.has_value ? opt.value : 0xFFFFFFFFFFFFFFFFull;
            for (int i = 0; i < 8; i++) re_put(e, (uint8_t)(v >> (8 * i)));
        }
        // process, dumbencoder.rs:64-85
        const uint8_t* src = in + it.in_off;
        uint32_t prev = 0;
        for (uint64_t k = 0; k < it.in_len; k++) {
            const uint32_t byte = __ldg(src + k);
            re_encode_bit(e, &is_match[k & 3], 0);
            uint16_t* probs = T + (prev >> 5) * 0x300;  // encode_literal, 124-139
            uint32_t result = 1;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t bit = (byte >> (7 - i)) & 1u;
                re_encode_bit(e, &probs[result], bit);
                result = (result << 1) ^ bit;
            }
            prev = byte;
        }
        // finish(input_len + 1), dumbencoder.rs:87-122: end marker for WriteToHeader(None) only.  input_len is the INDEX
        // of the last byte, so the marker's pos_state is in_len & 3 -- and 1 for an empty input, as in the reference.
        if (!opt.skip_size_field && !opt.has_value) {
            const uint64_t marker_pos = it.in_len ? it.in_len : 1;
            uint16_t fresh;
            re_encode_bit(e, &is_match[marker_pos & 3], 1);
            fresh = 0x400, re_encode_bit(e, &fresh, 0);                              // new distance
            for (int i = 0; i < 4; i++) fresh = 0x400, re_encode_bit(e, &fresh, 0);  // len = 0
            for (int i = 0; i < 36; i++) fresh = 0x400, re_encode_bit(e, &fresh, 1); // pos_slot 63 + 30 bits: 0xFFFFFFFF
        }
        for (int i = 0; i < 5; i++) re_write_low(e);  // RangeEncoder::finish, rangecoder.rs:54-61
        LzbEncResult r;
        r.code = e.pos <= e.cap ? LZB_OK : LZB_E_CAPACITY;
        r.pad = 0;
        r.out_len = e.pos;
        results[s] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// K6: decode of LZMA2 streams that consist of stored chunks only (LZB_ITEM_F_ALL_STORED: the framing scan walked them
// to their 0x00 terminator with every payload complete, and the output fits) -- "LZMA2 uncompressed chunks become
// straight memcpy".  One CTA per stream walks the 3-byte chunk headers (parse_uncompressed, lzma2.rs:195-229; a dict
// reset changes nothing for data that is only copied) and its 8 warps copy 8 KiB slices of each chunk.
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256)
    lzb_stored_decode_kernel(const LzbItem* __restrict__ items, const uint32_t* __restrict__ list,
                             const uint8_t* __restrict__ in, uint8_t* out, LzbResult* results) {
    const uint32_t idx = list[blockIdx.x];
    const LzbItem it = items[idx];
    const uint8_t* p = in + it.in_off;
    uint8_t* o = out + it.out_off;
    const int lane = threadIdx.x & 31;
    const uint32_t w0 = (threadIdx.x >> 5) * 8192u;
    uint64_t q = 0, opos = 0;
    uint32_t chunks = 0;
    for (;;) {
        const uint32_t status = __ldg(p + q);
        chunks++;
        q++;
        if (status == 0) break;
        const uint32_t n = ((__ldg(p + q) << 8) | __ldg(p + q + 1)) + 1;
        q += 2;
        if (w0 < n) warp_copy<true>(o + opos + w0, p + q + w0, n - w0 < 8192u ? n - w0 : 8192u, lane);
        opos += n;
        q += n;
    }
    if (threadIdx.x == 0) {
        LzbResult r;
        r.code = LZB_OK;
        r.chunks = chunks;
        r.a0 = r.a1 = 0;
        r.out_len = r.sink_len = opos;
        r.consumed = q;
        results[idx] = r;
    }
}
