// ASan/UBSan fuzz of dm_gunzip (inflates untrusted gzip / zlib bodies for the manifest hook), built natively
// with g++ by tests/test_native_host.py.  Seeds: real gzip / zlib streams produced by Python's zlib (stored,
// fixed-Huffman and dynamic-Huffman blocks) passed on the command line as files; each is mutated (bit flips,
// truncation, splices, random tails) 40 000 times.  Every outcome must be DM_OK (then the output must equal
// the seed's plain text or fail the CRC - i.e. rc DM_OK implies an intact checksum), DM_EINVAL or DM_ENOMEM;
// the sanitizers catch any over-read of the input or over-write of the exact-size output buffer.
#include "../../demodel_b200/csrc/gunzip.cc"

#include <cstdio>
#include <fstream>
#include <iterator>
#include <random>
#include <string>

int main(int argc, char **argv)
{
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 1; i < argc; ++i) {
        std::ifstream f(argv[i], std::ios::binary);
        seeds.emplace_back(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    }
    if (seeds.empty()) { fprintf(stderr, "usage: fuzz_gunzip seed.gz...\n"); return 2; }
    std::mt19937_64 rng(23);
    long ok = 0, bad = 0, small = 0;
    for (auto &sd : seeds) {                       // every seed must inflate as is
        std::vector<uint8_t> out(1 << 20);
        size_t n = 0;
        if (dm_gunzip(sd.data(), sd.size(), out.data(), out.size(), &n) != DM_OK) { fprintf(stderr, "seed rejected\n"); return 1; }
    }
    for (int it = 0; it < 40000 * (int)seeds.size(); ++it) {
        std::vector<uint8_t> s = seeds[it % seeds.size()];
        switch ((it / seeds.size()) % 5) {
        case 0: for (int k = 0, m = 1 + (int)(rng() % 4); k < m && !s.empty(); ++k) s[rng() % s.size()] ^= (uint8_t)(1u << (rng() % 8)); break;
        case 1: s.resize(rng() % (s.size() + 1)); break;
        case 2: if (s.size() > 20) { const size_t a = 10 + rng() % (s.size() - 18); s.insert(s.begin() + (long)a, (uint8_t)rng()); } break;
        case 3: for (size_t k = 10 + rng() % 30; k < s.size(); ++k) s[k] = (uint8_t)rng(); break;     // valid header, random deflate data
        default: { const auto &o = seeds[rng() % seeds.size()]; if (s.size() > 12 && o.size() > 12) { const size_t a = 10 + rng() % (s.size() - 10); s.resize(a); s.insert(s.end(), o.begin() + (long)(rng() % o.size()), o.end()); } } break;
        }
        const size_t cap = (it % 7 == 0) ? rng() % 64 : (size_t)1 << 17;
        uint8_t *in = new uint8_t[s.size() ? s.size() : 1];          // exact-size heap copies: over-reads / over-writes trip ASan
        if (!s.empty()) memcpy(in, s.data(), s.size());
        uint8_t *out = new uint8_t[cap ? cap : 1];
        size_t n = 0;
        const int rc = dm_gunzip(in, s.size(), out, cap, &n);
        if (rc == DM_OK) { ++ok; if (n > cap) { fprintf(stderr, "out_len beyond cap\n"); return 1; } }
        else if (rc == DM_EINVAL) ++bad;
        else if (rc == DM_ENOMEM) ++small;
        else { fprintf(stderr, "unexpected rc %d\n", rc); return 1; }
        delete[] in;
        delete[] out;
    }
    printf("gunzip fuzz ok: %ld inflated, %ld rejected, %ld too small\n", ok, bad, small);
    return 0;
}
