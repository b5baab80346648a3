// Internal interface between the engine (host C++) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dm {

// One unit of hashing work: advance one blob's SHA-256 state over a
// contiguous run of bytes in HBM.  Non-final jobs cover whole 64-byte blocks;
// a final job may end in a partial block, which the kernel pads on device
// (FIPS 180-4 §5.1.1) using total_len.
struct alignas(16) HashJob {
    const uint8_t *src;    // 16-byte aligned
    uint8_t *dst;          // CAS extent to copy into while hashing, or nullptr
    uint64_t nbytes;       // bytes of this run
    uint64_t total_len;    // whole-blob length (used when FINAL)
    uint32_t slot;         // index into the state table
    uint32_t flags;        // JOB_*
    uint32_t one;          // always 1: a multiplier the compiler cannot see (FMA-pipe adds)
    uint32_t pad_;
};
static_assert(sizeof(HashJob) == 48, "HashJob layout is shared with the kernels");

enum : uint32_t {
    JOB_INIT  = 1u,   // start from the FIPS 180-4 §5.3.3 initial hash value
    JOB_FINAL = 2u,   // apply padding and emit the digest words
};

// states/digests: uint32[slot][8], native word order (host serialises big-endian).
// `digests` may live in mapped pinned host memory.
// variant (tuning knob, every value is bit-identical): for the wide kernel
// fma + 3*style, for the deep kernel fma, where fma = how round additions are
// issued (0 = ptxas' choice, 1 = all on the FMA pipe, 2 = FMA pipe with the
// shortest e-chain) and style = main-loop shape (see sha256_wide_kernel).
cudaError_t launch_sha256_wide(const HashJob *jobs, uint32_t njobs, uint32_t *states,
                               uint32_t *digests, cudaStream_t stream, int variant);
cudaError_t launch_sha256_deep(const HashJob *jobs, uint32_t njobs, uint32_t *states,
                               uint32_t *digests, cudaStream_t stream, int variant);
constexpr int kDefaultWideVariant = 7;
constexpr int kDefaultDeepVariant = 0;

// Live streams below this count go to the warp-per-stream (deep) kernel,
// above it to the lane-per-stream (wide) kernel; see DESIGN.md §4.
constexpr uint32_t kDeepWideCrossover = 1024;

cudaError_t launch_synth_fill(uint64_t seed, uint64_t blob, uint64_t byte_off, void *dst,
                              size_t len, cudaStream_t stream);
// dev_offsets / dev_lengths: n entries each, offsets ascending and 16-byte aligned;
// span = offsets[n-1] + lengths[n-1] - offsets[0].
cudaError_t launch_synth_fill_many(uint64_t seed, uint64_t first_blob, void *base,
                                   const uint64_t *dev_offsets, const uint64_t *dev_lengths, uint32_t n,
                                   uint64_t first_off, uint64_t span, cudaStream_t stream);

}  // namespace dm
