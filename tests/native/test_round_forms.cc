// Host-side proof that every round form in demodel_b200/csrc/sha256_round.cuh computes SHA-256.
// The header compiles as plain C++ here (the PTX one-instruction forms are replaced by the portable
// expressions next to them), so the ORDER of operations of each form - which is what differs between
// them - is exactly what the kernels run.  Checked against the CPU oracle (test infrastructure).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../demodel_b200/csrc/sha256_round.cuh"

extern "C" void dmo_sha256(const void *data, size_t len, uint8_t out[32]);

namespace {

template <int kFma, int t>
struct Rounds {
    static void run(uint32_t (&v)[8], const uint32_t (&kw)[64], const dm::FmaK &k)
    {
        dm::sha_round<kFma, t>(v, kw[t], k);
        if constexpr (t < 63) Rounds<kFma, t + 1>::run(v, kw, k);
    }
};

template <int kFma>
void compress(uint32_t (&s)[8], const uint8_t *blk, const dm::FmaK &k)
{
    constexpr uint32_t K[64] = {DM_K256_TABLE};
    uint32_t w[64], kw[64], v[8];
    for (int i = 0; i < 16; ++i) { uint32_t x; memcpy(&x, blk + 4 * i, 4); w[i] = dm::bswap32(x); }
    for (int i = 16; i < 64; ++i)
        w[i] = dm::addf<kFma>(dm::addf<kFma>(dm::addf<kFma>(dm::small_sigma1(w[i - 2]), w[i - 7], k), dm::small_sigma0(w[i - 15]), k), w[i - 16], k);
    for (int i = 0; i < 64; ++i) kw[i] = dm::addf<kFma>(w[i], K[i], k);
    for (int i = 0; i < 8; ++i) v[i] = s[i];
    Rounds<kFma, 0>::run(v, kw, k);
    for (int i = 0; i < 8; ++i) s[i] += v[i];
}

template <int kFma>
void sha256(const uint8_t *m, size_t n, uint8_t out[32])
{
    const dm::FmaK k = {1u, 0xffffffffu, {0u, 0u}};
    uint32_t s[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    size_t off = 0;
    for (; off + 64 <= n; off += 64) compress<kFma>(s, m + off, k);
    uint8_t tail[128] = {0};
    const size_t r = n - off;
    if (r) memcpy(tail, m + off, r);
    tail[r] = 0x80;
    const size_t tl = r + 9 <= 64 ? 64 : 128;
    const uint64_t bits = (uint64_t)n * 8;
    for (int i = 0; i < 8; ++i) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    compress<kFma>(s, tail, k);
    if (tl == 128) compress<kFma>(s, tail + 64, k);
    for (int i = 0; i < 8; ++i) { const uint32_t x = dm::bswap32(s[i]); memcpy(out + 4 * i, &x, 4); }
}

template <int kFma>
int check(const char *name)
{
    uint64_t x = 0x9e3779b97f4a7c15ull + kFma;
    auto next = [&] { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    int bad = 0, n = 0;
    std::vector<size_t> lens;
    for (size_t l = 0; l <= 300; ++l) lens.push_back(l);
    for (size_t l : {1000u, 4096u, 65536u + 1u, 1000003u}) lens.push_back(l);
    for (size_t l : lens) {
        std::vector<uint8_t> m(l);
        for (auto &b : m) b = (uint8_t)next();
        uint8_t a[32], b[32];
        sha256<kFma>(m.data(), l, a);
        dmo_sha256(m.data(), l, b);
        if (memcmp(a, b, 32)) { if (!bad) fprintf(stderr, "%s: first mismatch at length %zu\n", name, l); ++bad; }
        ++n;
    }
    printf("%-28s %d messages, %d mismatches\n", name, n, bad);
    return bad;
}

}  // namespace

int main()
{
    int bad = 0;
    bad += check<0>("kFma 0 (ptxas ordering)");
    bad += check<1>("kFma 1 (IMAD additions)");
    bad += check<4>("kFma 4 (short chain)");
    bad += check<5>("kFma 5 (short chain, all IMAD)");
    bad += check<6>("kFma 6 (e' IADD3, a' IMAD)");
    bad += check<7>("kFma 7 (e' IMAD, a' IADD3)");
    if (bad) { printf("ROUND FORMS FAILED\n"); return 1; }
    printf("ROUND FORMS OK\n");
    return 0;
}
