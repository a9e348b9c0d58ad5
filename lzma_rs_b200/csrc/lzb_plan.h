// lzb_plan.h -- host-side planning around the decode kernels: header parsing, LZMA2 framing scan, the XZ
// container walk and status rendering.  Pure C++ (no CUDA); the device work is reached through `Executor`.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "lzb_types.h"

namespace lzb {

struct CrcRange {
    uint64_t off, len;  // into the output blob
};

// Where work items run.  The shipped library has exactly one implementation (CUDA, lzb_host.cu); the
// CPU test harness under tests/host_emulation provides a second one to check this file's logic without a GPU.
class Executor {
   public:
    virtual ~Executor() {}
    // Decode items[0..n) (offsets relative to the bound input / output blobs).  Returns LZB_RC_*.
    // stored_bytes: output bytes the framing scan saw in stored (uncompressed) LZMA2 chunks; selects kernel variants,
    // never changes results
    virtual int decode(const LzbItem* items, uint32_t n, uint32_t max_lclp, uint64_t stored_bytes, LzbResult* results) = 0;
    // CRC-32 and CRC-64 of output-blob ranges.
    virtual int crc(const CrcRange* ranges, uint32_t n, uint32_t* crc32, uint64_t* crc64) = 0;
    // Device scratch of `bytes` bytes, returned as an offset in output-blob coordinates (valid until the executor dies):
    // intermediate results of chained .xz filters.  Returns LZB_RC_*.
    virtual int scratch(uint64_t bytes, uint64_t* out_off) = 0;
    // Copies output-blob bytes [off, off+len) to host memory (the framing scan of an intermediate result).
    virtual int read_out(uint64_t off, uint64_t len, uint8_t* dst) = 0;
};

// ---- scans (host copies of what K2 does on the device) ----
struct Lzma2Scan {
    uint64_t unpacked = 0;  // sum of chunk unpacked sizes
    uint64_t packed = 0;    // bytes up to and including the 0x00 control byte
    uint32_t max_lclp = 0;
    bool well_formed = false;  // walk ended at a 0x00 control byte with valid framing
    uint64_t stored = 0;       // bytes of `unpacked` that sit in stored (uncompressed) chunks
};
Lzma2Scan scan_lzma2(const uint8_t* p, uint64_t len);

// .lzma header -> item (kind LZB_ITEM_LZMA or LZB_ITEM_PRESET).  in_off/in_len are relative to `base_off`.
void plan_lzma(const uint8_t* p, uint64_t len, uint64_t base_off, const lzb_options* opt, LzbItem* it, LzbScan* sc);
void plan_lzma2(const uint8_t* p, uint64_t len, uint64_t base_off, LzbItem* it, LzbScan* sc);
bool lenient_eof(int fmt, const lzb_options* opt, const LzbResult* r);  // Options::allow_incomplete applies

// Per-stream outcome of a batch call.
struct StreamOut {
    lzb_status st{};
    uint64_t out_len = 0;
    uint64_t consumed = 0;
};

void status_from_result(const LzbResult& r, lzb_status* st);

// Decodes n .xz files (xz.rs:18-94): container walk on the host, one work item per block on the executor,
// block checks from executor CRCs.  File i = in[in_off[i], in_off[i+1]); its output goes to
// [out_off[i], out_off[i+1]) of the executor's output blob.  Returns LZB_RC_*.
int decode_xz_batch(Executor& ex, const uint8_t* in, const uint64_t* in_off, uint32_t n, const uint64_t* out_off,
                    StreamOut* outs);

// All three formats over a batch of host-resident inputs: plans work items, runs them on `ex`, maps results.
int decode_batch(Executor& ex, int fmt, const lzb_options* opt, const uint8_t* in, const uint64_t* in_off, uint32_t n,
                 const uint64_t* out_off, StreamOut* outs);
uint64_t scan_capacity(int fmt, const lzb_options* opt, const uint8_t* p, uint64_t len);

// Capacity lzb_decode_batch needs for one .xz file (sum of block sizes; header sizes if present, else scan).
uint64_t scan_xz_capacity(const uint8_t* p, uint64_t len);

uint32_t crc32_host(const uint8_t* p, size_t n);  // container CRCs (headers, index, footer): a few bytes each

}  // namespace lzb
