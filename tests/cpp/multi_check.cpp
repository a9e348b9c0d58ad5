// Client of the multi-device entry points of include/lzma_b200.h, compiled with g++ (no CUDA headers):
//   multi_check <devices: "all" | "0,0,1"> <raw LZMA2 stream file> <expected plaintext file> <copies>
// builds a host batch of <copies> streams (the file's stream, every third one truncated to provoke an error status),
// decodes it with lzb_decode_batch_multi and checks every output and status.  Exit 0 = all good; prints the split.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "lzma_b200.h"

static std::vector<uint8_t> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
    if (argc != 5) return 2;
    std::vector<int> devs;
    if (std::string(argv[1]) != "all") {
        char* s = argv[1];
        for (char* tok = strtok(s, ","); tok; tok = strtok(nullptr, ",")) devs.push_back(atoi(tok));
    }
    const std::vector<uint8_t> stream = slurp(argv[2]), plain = slurp(argv[3]);
    const uint32_t n = (uint32_t)atoi(argv[4]);
    lzb_multi* m = nullptr;
    int rc = lzb_create_multi(&m, devs.empty() ? nullptr : devs.data(), (int)devs.size());
    if (rc != LZB_RC_OK) {
        fprintf(stderr, "lzb_create_multi failed: %d\n", rc);
        return 3;
    }
    const int nd = lzb_multi_device_count(m);
    std::vector<uint64_t> in_off(n + 1, 0), out_off(n + 1, 0), out_len(n), consumed(n);
    std::vector<uint8_t> blob;
    for (uint32_t i = 0; i < n; i++) {
        const size_t len = i % 3 == 2 ? stream.size() / 2 : stream.size();
        blob.insert(blob.end(), stream.begin(), stream.begin() + len);
        in_off[i + 1] = blob.size();
        out_off[i + 1] = out_off[i] + ((plain.size() + 15) & ~(size_t)15);
    }
    blob.resize(blob.size() + 16);
    std::vector<uint8_t> out(out_off[n] + 16, 0xEE);
    std::vector<lzb_status> st(n);
    std::vector<uint32_t> split(nd + 1);
    rc = lzb_decode_batch_multi(m, LZB_FMT_LZMA2, nullptr, blob.data(), in_off.data(), n, out.data(), out_off.data(),
                                out_len.data(), consumed.data(), st.data(), split.data());
    if (rc != LZB_RC_OK) {
        fprintf(stderr, "lzb_decode_batch_multi failed: %d (%s)\n", rc, lzb_multi_last_error(m));
        return 3;
    }
    printf("devices %d split", nd);
    for (int k = 0; k <= nd; k++) printf(" %u", split[k]);
    printf("\n");
    int bad = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (i % 3 == 2) {  // truncated: UnexpectedEof, the complete part of the output is still there
            char msg[256];
            lzb_format_error(&st[i], msg, sizeof msg);
            if (st[i].code == LZB_OK || strcmp(msg, "io error: failed to fill whole buffer") != 0) bad++;
            if (out_len[i] > plain.size() || memcmp(out.data() + out_off[i], plain.data(), out_len[i]) != 0) bad++;
        } else {
            if (st[i].code != LZB_OK || out_len[i] != plain.size() || consumed[i] != stream.size() ||
                memcmp(out.data() + out_off[i], plain.data(), plain.size()) != 0)
                bad++;
        }
    }
    lzb_destroy_multi(m);
    if (bad) fprintf(stderr, "%d streams wrong\n", bad);
    return bad ? 1 : 0;
}
