// Exercises the C++ mirror of the reference API (lzma_rs_b200/host/lzma_rs.hpp) end to end on a GPU box:
//   host_check <rt-lzma|rt-lzma2|rt-xz> <plain file> <output file>     -> compress then decompress (tests/lzma.rs:16-28 style)
//   host_check <lzma|lzma2|xz|rawlzma> <input file> <output file>   -> exit 0 and writes the decoded bytes, or prints the
//   reference-format error string to stderr and exits 3 (partial output is still written).
#include <fstream>
#include <iostream>
#include <sstream>
#include <iterator>
#include <string>
#include <vector>

#include "lzma_rs_b200/host/lzma_rs.hpp"

int main(int argc, char** argv) {
    if (argc != 4) return 2;
    const std::string fmt = argv[1];
    std::ifstream in(argv[2], std::ios::binary);
    std::ofstream out(argv[3], std::ios::binary);
    try {
        if (fmt.rfind("rt-", 0) == 0) {  // round trip through the compress side
            std::stringstream packed;
            if (fmt == "rt-lzma") lzma_rs::lzma_compress(in, packed);
            else if (fmt == "rt-lzma2") lzma_rs::lzma2_compress(in, packed);
            else lzma_rs::xz_compress(in, packed);
            if (fmt == "rt-lzma") lzma_rs::lzma_decompress(packed, out);
            else if (fmt == "rt-lzma2") lzma_rs::lzma2_decompress(packed, out);
            else lzma_rs::xz_decompress(packed, out);
            return 0;
        }
        if (fmt == "batch-lzma2") {  // batch form over two contexts (the same GPU twice): 7 copies, one of them truncated
            std::vector<char> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
            std::vector<lzma_rs::batch::Input> ins(7, {reinterpret_cast<const uint8_t*>(buf.data()), buf.size()});
            ins[3].size = buf.size() / 2;
            lzma_rs::batch::set_devices({0, 0});
            auto res = lzma_rs::batch::lzma2_decompress_batch(ins);
            for (size_t i = 0; i < res.size(); i++) {
                if (i == 3) {
                    if (res[i].ok() || res[i].error != "io error: failed to fill whole buffer") return 5;
                } else if (!res[i].ok() || res[i].consumed != buf.size() || res[i].data != res[0].data) {
                    return 6;
                }
            }
            out.write(reinterpret_cast<const char*>(res[0].data.data()), (std::streamsize)res[0].data.size());
            return 0;
        }
        if (fmt == "lzma") lzma_rs::lzma_decompress(in, out);
        else if (fmt == "lzma2") lzma_rs::lzma2_decompress(in, out);
        else if (fmt == "rawlzma") {  // decompress::raw: split the 13-byte .lzma header off, decode the headerless payload
            namespace raw = lzma_rs::decompress::raw;
            unsigned char h[13];
            in.read(reinterpret_cast<char*>(h), 13);
            uint32_t dict = 0;
            uint64_t size = 0;
            for (int k = 3; k >= 0; k--) dict = (dict << 8) | h[1 + k];
            for (int k = 7; k >= 0; k--) size = (size << 8) | h[5 + k];
            raw::LzmaParams params{{h[0] % 9u, (h[0] / 9u) % 5u, h[0] / 45u}, dict < 0x1000 ? 0x1000 : dict,
                                   size == ~0ull ? std::nullopt : std::optional<uint64_t>(size)};
            raw::LzmaDecoder dec(params, std::nullopt);
            dec.decompress(in, out);
            // decoding again without reset() continues from the carried DecoderState like the reference (lzma.rs:597-633);
            // after reset() the same payload decodes to the same bytes again
            in.clear();
            in.seekg(13);
            std::ostringstream again;
            dec.reset();
            dec.decompress(in, again);
            std::ifstream first(argv[3], std::ios::binary);
            out.flush();
            std::string a((std::istreambuf_iterator<char>(first)), std::istreambuf_iterator<char>());
            if (again.str() != a) return 4;
        }
        else lzma_rs::xz_decompress(in, out);
    } catch (const lzma_rs::error::Error& e) {
        std::cerr << e.what();
        return 3;
    }
    return 0;
}
