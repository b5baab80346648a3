// TEST INFRASTRUCTURE ONLY — see fake_cuda/cuda_runtime.h.  Synchronous fake runtime + the three
// SHA-256 "kernel" launchers executed on the CPU by the oracle (oracle/sha256_oracle.c), honouring
// exactly the HashJob contract the real kernels implement (INIT / FINAL flags, state table, fused copy).
#include "fake_cuda/cuda_runtime.h"
#include "../../demodel_b200/csrc/sha256_kernels.cuh"
#include "../../demodel_b200/csrc/blobgen.h"

#include <cstdlib>
#include <cstring>

extern "C" {
#include "../../oracle/sha256_oracle.c"
}

extern "C" {
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof *p); p->major = 10; p->multiProcessorCount = 148; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = 1ull << 30; return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "fake cuda error"; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void **dev, void *host, unsigned) { *dev = host; return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.001f; return cudaSuccess; }
}

namespace dm {

// One HashJob, the way every real kernel shape treats it.
static void run_job(const HashJob &jb, uint32_t *states, uint32_t *digests)
{
    dmo_sha256_ctx c;
    if (jb.flags & JOB_INIT) dmo_sha256_init(&c);
    else {
        memcpy(c.h, states + 8ull * jb.slot, 32);
        c.nbuf = 0;
        c.nbytes = 0;
    }
    if (jb.nbytes) {
        if (jb.dst) memcpy(jb.dst, jb.src, jb.nbytes);
        if (jb.flags & JOB_FINAL) dmo_sha256_update(&c, jb.src, jb.nbytes);
        else dmo_sha256_update(&c, jb.src, jb.nbytes & ~63ull);      // non-final jobs are whole blocks by contract
    }
    if (jb.flags & JOB_FINAL) {
        c.nbytes = jb.total_len;                                     // padding carries the whole-blob length
        uint8_t out[32];
        dmo_sha256_final(&c, out);
        uint32_t w[8];
        for (int i = 0; i < 8; ++i) w[i] = ((uint32_t)out[4 * i] << 24) | ((uint32_t)out[4 * i + 1] << 16) | ((uint32_t)out[4 * i + 2] << 8) | out[4 * i + 3];
        memcpy(states + 8ull * jb.slot, w, 32);
        memcpy(digests + 8ull * jb.slot, w, 32);
    } else {
        memcpy(states + 8ull * jb.slot, c.h, 32);
    }
}

static cudaError_t run_all(const HashJob *jobs, uint32_t n, uint32_t *states, uint32_t *digests)
{
    for (uint32_t i = 0; i < n; ++i) run_job(jobs[i], states, digests);
    return cudaSuccess;
}
cudaError_t launch_sha256_wide(const HashJob *j, uint32_t n, uint32_t *s, uint32_t *d, cudaStream_t, int) { return run_all(j, n, s, d); }
cudaError_t launch_sha256_deep(const HashJob *j, uint32_t n, uint32_t *s, uint32_t *d, cudaStream_t, int) { return run_all(j, n, s, d); }
cudaError_t launch_sha256_group(const HashJob *j, uint32_t n, uint32_t *s, uint32_t *d, cudaStream_t, int) { return run_all(j, n, s, d); }

cudaError_t launch_synth_fill(uint64_t seed, uint64_t blob, uint64_t off, void *dst, size_t len, cudaStream_t)
{
    dmo_blob_fill(seed, blob, off, dst, len);
    return cudaSuccess;
}
cudaError_t launch_synth_fill_many(uint64_t seed, uint64_t first, void *base, const uint64_t *offs, const uint64_t *lens,
                                   uint32_t n, uint64_t, uint64_t, cudaStream_t)
{
    for (uint32_t i = 0; i < n; ++i) dmo_blob_fill(seed, first + i, 0, static_cast<uint8_t *>(base) + offs[i], lens[i]);
    return cudaSuccess;
}

}  // namespace dm
