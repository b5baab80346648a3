// Synthetic blob bytes: a counter-based generator so a blob
// can be produced wherever it is consumed (host ring or HBM) and its digest
// cross-checked without shipping bytes.  Shared by host and device code.
//
// Byte j of blob b under seed s is byte (j & 7), little-endian, of
//   mix64( mix64(s + PHI*(b+1)) + GAMMA*((j>>3)+1) )
// where mix64 is the SplitMix64 output finaliser.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD static inline
#endif

DM_HD uint64_t dm_mix64(uint64_t z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

DM_HD uint64_t dm_blob_key(uint64_t seed, uint64_t blob)
{
    return dm_mix64(seed + 0x9E3779B97F4A7C15ull * (blob + 1));
}

DM_HD uint64_t dm_blob_word_k(uint64_t key, uint64_t word_index)
{
    return dm_mix64(key + 0xD1B54A32D192ED03ull * (word_index + 1));
}
