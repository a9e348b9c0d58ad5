// lzb_encode.h -- records shared by the host side and the encode kernels (lzb_encode_kernels.cu).
#pragma once
#include <stdint.h>

#include "lzma_b200.h"

#if defined(__CUDACC__)
#define LZB_HD __host__ __device__ inline
#else
#define LZB_HD static inline
#endif

struct LzbEncItem {
    uint64_t in_off, in_len;    // plaintext of this stream in the input blob
    uint64_t out_off, out_cap;  // where its encoding goes in the output blob
};
struct LzbEncResult {
    int32_t code;  // LZB_OK / LZB_E_CAPACITY
    uint32_t pad;
    uint64_t out_len;  // bytes of the encoding (the size needed when code == LZB_E_CAPACITY)
};
struct LzbXzHead {
    uint8_t b[24];  // stream header (xz.rs:31-45) + block header (xz.rs:78-97): the same bytes for every stream
};

#define LZB_ENC_LANES 16                        // streams per CTA of K5 (one thread each)
#define LZB_ENC_TABLE_U16 (8 * 0x300 + 4)       // literal_probs[8][0x300] + is_match[4], dumbencoder.rs:10-12

// Exact encoded size of the stored-chunk formats (LZMA2: 3 bytes per 64 KiB chunk + terminator, lzma2.rs:4-26;
// XZ: 12 + 12 header bytes, block padding, index (1 + 1 + two multibyte integers, padded, + CRC32), 12 footer bytes).
LZB_HD uint64_t lzb_multibyte_len(uint64_t v) {
    uint64_t n = 1;
    while (v >>= 7) n++;
    return n;
}
LZB_HD uint64_t lzb_encode_exact(int fmt, uint64_t in_len) {
    const uint64_t l2 = in_len + 3 * ((in_len + 0xFFFFull) >> 16) + 1;
    if (fmt == LZB_FMT_LZMA2) return l2;
    const uint64_t unpadded = 12 + l2;
    uint64_t index = 2 + lzb_multibyte_len(unpadded) + lzb_multibyte_len(in_len);
    index = ((index + 3) & ~3ull) + 4;
    return 12 + ((unpadded + 3) & ~3ull) + index + 12;
}
