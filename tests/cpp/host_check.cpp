// Exercises the C++ mirror of the reference API (lzma_rs_b200/host/lzma_rs.hpp) end to end on a GPU box:
//   host_check <lzma|lzma2|xz|rawlzma> <input file> <output file>   -> exit 0 and writes the decoded bytes, or prints the
//   reference-format error string to stderr and exits 3 (partial output is still written).
#include <fstream>
#include <iostream>
#include <string>

#include "lzma_rs_b200/host/lzma_rs.hpp"

int main(int argc, char** argv) {
    if (argc != 4) return 2;
    const std::string fmt = argv[1];
    std::ifstream in(argv[2], std::ios::binary);
    std::ofstream out(argv[3], std::ios::binary);
    try {
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
            try {
                dec.decompress(in, out);
                return 4;  // must not decode again without reset()
            } catch (const std::logic_error&) {
            }
        }
        else lzma_rs::xz_decompress(in, out);
    } catch (const lzma_rs::error::Error& e) {
        std::cerr << e.what();
        return 3;
    }
    return 0;
}
