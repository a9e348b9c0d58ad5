// Exercises the C++ mirror of the reference API (lzma_rs_b200/host/lzma_rs.hpp) end to end on a GPU box:
//   host_check <lzma|lzma2|xz> <input file> <output file>   -> exit 0 and writes the decoded bytes, or prints the
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
        else lzma_rs::xz_decompress(in, out);
    } catch (const lzma_rs::error::Error& e) {
        std::cerr << e.what();
        return 3;
    }
    return 0;
}
