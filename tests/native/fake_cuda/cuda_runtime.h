// TEST INFRASTRUCTURE ONLY — a fake "CUDA runtime" (asynchronous: one worker thread per stream) so that the engine's host
// logic (demodel_b200/csrc/engine_*.cu: ring, pump thread, CAS, ranges, followers, disk tier) can be
// compiled with plain g++ and soaked under ThreadSanitizer / ASan on a box with no GPU.
// Device memory is host memory; copies and "kernels" run later on the stream's worker thread in FIFO order;
// events carry record generations (query / synchronize / stream-wait behave like CUDA's); the SHA-256
// "kernels" are executed by the CPU oracle (tests/native/fake_cuda.cc).
// Never linked into libdemodel_b200.so; built only by tests/test_native_host.py into a temp dir.
#pragma once
#include <cstddef>
#include <cstdint>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorUnknown = 999 };
typedef struct fakeStream *cudaStream_t;
typedef struct fakeEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; };

extern "C" {
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int dev);
cudaError_t cudaDeviceGetPCIBusId(char *buf, int len, int dev);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b);
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError(void);
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned flags);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaHostGetDevicePointer(void **dev, void *host, unsigned flags);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventQuery(cudaEvent_t e);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
}

// the typed convenience overloads the real header has
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)(void *)p, n); }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned f) { return cudaHostAlloc((void **)(void *)p, n, f); }
