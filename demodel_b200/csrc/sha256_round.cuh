// The SHA-256 round arithmetic shared by every kernel shape, in a header of its own so that the SAME
// source also compiles as plain C++ (g++, no CUDA): tests/native/test_round_forms.cc runs every round
// form below against the CPU oracle, which is how a new instruction ordering is proven bit-exact before
// it ever reaches a GPU.  Under nvcc the functions are __device__ and use one-instruction PTX forms;
// under a host compiler the same expressions are written in portable C.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define DM_RD __device__ __forceinline__
#else
#define DM_RD static inline
#endif

namespace dm {
namespace {

// ---------------------------------------------------------------------------
// FIPS 180-4 §4.1.2 functions as single SASS ops.
// ---------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t rotr(uint32_t x, uint32_t n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t f_ch(uint32_t e, uint32_t f, uint32_t g)
{
    uint32_t d;   // (e & f) ^ (~e & g)
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(d) : "r"(e), "r"(f), "r"(g));
    return d;
}
__device__ __forceinline__ uint32_t f_maj(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;   // (a & b) ^ (a & c) ^ (b & c)
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// a*m + b on the FMA pipe (IMAD); m is a runtime value so ptxas cannot fold it back into IADD3
__device__ __forceinline__ uint32_t mad32(uint32_t a, uint32_t m, uint32_t b)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(b));
    return d;
}
// a + b + c as two PTX adds the front end cannot reassociate: ptxas fuses them into ONE IADD3 (ALU pipe)
__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
#else
static inline uint32_t rotr(uint32_t x, uint32_t n) { return (x >> n) | (x << ((32u - n) & 31u)); }
static inline uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
static inline uint32_t f_ch(uint32_t e, uint32_t f, uint32_t g) { return (e & f) ^ (~e & g); }
static inline uint32_t f_maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) ^ (a & c) ^ (b & c); }
static inline uint32_t bswap32(uint32_t x) { return __builtin_bswap32(x); }
static inline uint32_t mad32(uint32_t a, uint32_t m, uint32_t b) { return a * m + b; }
static inline uint32_t add3(uint32_t a, uint32_t b, uint32_t c) { return a + b + c; }
#endif
DM_RD uint32_t big_sigma0(uint32_t x) { return xor3(rotr(x, 2), rotr(x, 13), rotr(x, 22)); }
DM_RD uint32_t big_sigma1(uint32_t x) { return xor3(rotr(x, 6), rotr(x, 11), rotr(x, 25)); }
DM_RD uint32_t small_sigma0(uint32_t x) { return xor3(rotr(x, 7), rotr(x, 18), x >> 3); }
DM_RD uint32_t small_sigma1(uint32_t x) { return xor3(rotr(x, 17), rotr(x, 19), x >> 10); }

// FIPS 180-4 §4.2.2.  Indexed only with compile-time constants inside fully
// unrolled loops, so every use folds to an instruction immediate.
#define DM_K256_TABLE                                                                        \
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, \
    0xab1c5ed5u, 0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, \
    0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, \
    0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, \
    0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, \
    0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, \
    0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, \
    0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, \
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, \
    0xc67178f2u

// Runtime constants for FMA-pipe tricks (kernel argument; values fixed by the launcher).
struct FmaK {
    uint32_t one;                         // 1
    uint32_t neg;                         // 0xffffffff: x * neg + y == y - x on the FMA pipe
    uint32_t pad[2];
};

// Addition on the FMA pipe.  The integer ALU pipe (SHF/LOP3/IADD3) is the
// bottleneck of SHA-256 on this part: ~84% of the round instructions can only
// run there.  IMAD runs on the other (FMA) pipe, so `a*one + b` with a
// *runtime* one (a kernel argument: ptxas cannot fold it back into IADD3)
// moves the additions off the critical pipe.  kFma = 0 leaves the choice to
// ptxas, 1 forces every round/schedule addition onto the FMA pipe with the
// multiplier as a constant-bank operand, 2 does the same with the multiplier
// held in a register (loaded from the job record).  kFma = 4 is the
// "short chain" round below; its other additions are left to ptxas.
template <int kFma>
DM_RD uint32_t addf(uint32_t a, uint32_t b, const FmaK &k)
{
    if constexpr (kFma == 0 || kFma >= 4) {      // 4..7: round forms of their own; other additions left to ptxas
        (void)k;
        return a + b;
    } else {
        return mad32(a, k.one, b);
    }
}

// One round, FIPS 180-4 §6.2.2 step 3, with the a..h rotation done by
// renaming: v[] is indexed modulo 8 by the (compile-time) round number.
//   T1 = h + S1(e) + Ch(e,f,g) + (K+W);  d += T1;  h = T1 + S0(a) + Maj(a,b,c)
//
// kFma = 4, the short chain.  When ONE warp runs the rounds of one stream (deep and group kernels)
// nothing hides latency: the round time is the longer of the warp's ALU-pipe issue time (2 cycles per
// SHF/LOP3/IADD3) and the dependent chain e -> S1(e) -> ... -> e'.  ptxas' own ordering (kFma = 0) puts
// three dependent additions after S1 (IADD3 -> IMAD -> IMAD, two of them crossing pipes: 4+4+5+4+5
// cycles after the last SHF), 29 cycles per round.  Here everything that does not depend on e is summed
// first, on the FMA pipe (dx = d + h + K+W, md = Maj - d), so that each chain ends in ONE IADD3:
//     e' = dx + Ch(e,f,g) + S1(e)            a' = e' + S0(a) + md        (= T1 + T2, since e' - d = T1)
// i.e. SHF -> LOP3 -> IADD3 per round on both chains (4+4+4 after the third SHF = 16 cycles) with
// 12 ALU-pipe instructions (24 cycles of issue) and 3 IMADs.  tools/sass_sched.py on the SASS: 24.5
// cycles per round against 29.0 for kFma = 0 (whose prediction matches the measured 16.57 ms launch).
template <int kFma, int t>
DM_RD void sha_round(uint32_t (&v)[8], uint32_t kw, const FmaK &k)
{
    constexpr int ia = (0 - t) & 7, ib = (1 - t) & 7, ic = (2 - t) & 7, id = (3 - t) & 7;
    constexpr int ie = (4 - t) & 7, jf = (5 - t) & 7, ig = (6 - t) & 7, ih = (7 - t) & 7;
    if constexpr (kFma == 0) {
        const uint32_t t1 = v[ih] + big_sigma1(v[ie]) + f_ch(v[ie], v[jf], v[ig]) + kw;
        const uint32_t t2 = big_sigma0(v[ia]) + f_maj(v[ia], v[ib], v[ic]);
        v[id] += t1;
        v[ih] = t1 + t2;
    } else if constexpr (kFma == 4) {
        const uint32_t dx = mad32(mad32(v[ih], k.one, kw), k.one, v[id]);
        const uint32_t md = mad32(v[id], k.neg, f_maj(v[ia], v[ib], v[ic]));
        const uint32_t en = add3(dx, f_ch(v[ie], v[jf], v[ig]), big_sigma1(v[ie]));
        v[ih] = add3(md, big_sigma0(v[ia]), en);
        v[id] = en;
    } else if constexpr (kFma == 5) {
        // 10 ALU-pipe instructions (6 SHF + 4 LOP3), every addition an IMAD: 20 cycles of ALU issue per round.
        // Each chain still ends one instruction after its LOP3:  e' = S1 + (Ch + dx),  a' = S0 + (e' + md).
        const uint32_t dx = mad32(mad32(v[ih], k.one, kw), k.one, v[id]);
        const uint32_t md = mad32(v[id], k.neg, f_maj(v[ia], v[ib], v[ic]));
        const uint32_t cx = mad32(f_ch(v[ie], v[jf], v[ig]), k.one, dx);
        const uint32_t en = mad32(big_sigma1(v[ie]), k.one, cx);
        const uint32_t tm = mad32(en, k.one, md);
        v[ih] = mad32(big_sigma0(v[ia]), k.one, tm);
        v[id] = en;
    } else if constexpr (kFma == 6) {
        // e' on the ALU pipe (one IADD3, shortest e-chain), a' on the FMA pipe: 11 ALU-pipe instructions.
        const uint32_t dx = mad32(mad32(v[ih], k.one, kw), k.one, v[id]);
        const uint32_t md = mad32(v[id], k.neg, f_maj(v[ia], v[ib], v[ic]));
        const uint32_t en = add3(dx, f_ch(v[ie], v[jf], v[ig]), big_sigma1(v[ie]));
        const uint32_t tm = mad32(en, k.one, md);
        v[ih] = mad32(big_sigma0(v[ia]), k.one, tm);
        v[id] = en;
    } else if constexpr (kFma == 7) {
        // e' on the FMA pipe, a' one IADD3: 11 ALU-pipe instructions.
        const uint32_t dx = mad32(mad32(v[ih], k.one, kw), k.one, v[id]);
        const uint32_t md = mad32(v[id], k.neg, f_maj(v[ia], v[ib], v[ic]));
        const uint32_t cx = mad32(f_ch(v[ie], v[jf], v[ig]), k.one, dx);
        const uint32_t en = mad32(big_sigma1(v[ie]), k.one, cx);
        v[ih] = add3(md, big_sigma0(v[ia]), en);
        v[id] = en;
    } else {
        const uint32_t x = addf<1>(v[ih], kw, k);
        const uint32_t y = addf<1>(x, f_ch(v[ie], v[jf], v[ig]), k);
        const uint32_t t1 = addf<1>(y, big_sigma1(v[ie]), k);
        const uint32_t t2 = addf<1>(big_sigma0(v[ia]), f_maj(v[ia], v[ib], v[ic]), k);
        v[id] = addf<1>(v[id], t1, k);
        v[ih] = addf<1>(t1, t2, k);
    }
}

}  // namespace
}  // namespace dm
