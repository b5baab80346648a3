// dm_proxy_drive / dm_proxy_serve: the per-connection work of the proxy,
// restated as C++ threads over the hook wrappers in proxy_hooks.hpp.
//
// In the reference, net/http runs one goroutine per client connection
// (/root/reference/cmd/demodel/start.go:210-215) and goproxy reads the
// response body on that goroutine in 32 KiB pieces.  Here one OS thread plays
// one such goroutine: it takes the next blob, wraps its "upstream body" (a
// region of caller-owned host memory) in a BodyTee, and pumps Read() until
// EOF, exactly the loop goproxy's copy would run.
#include "proxy_hooks.hpp"

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

namespace {

struct MemUpstream : dm::Upstream {
    const uint8_t *p;
    uint64_t left;
    MemUpstream(const uint8_t *base, uint64_t n) : p(base), left(n) {}
    long Read(void *dst, size_t n) override
    {
        if (n > left) n = (size_t)left;
        memcpy(dst, p, n);
        p += n; left -= n;
        return (long)n;
    }
};

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" int dm_proxy_drive(dm_engine *e, const void *host_base, const uint64_t *offsets, uint32_t n,
                              const uint8_t *expect, size_t chunk, uint32_t concurrency, int nthreads,
                              int zero_copy, uint8_t *digests_out, uint8_t *matched_out, double *seconds)
{
    if (!e || !offsets || (!host_base && n)) return DM_EINVAL;
    if (chunk == 0) chunk = 32768;                    // io.Copy's buffer
    if (concurrency == 0) concurrency = 1;
    uint32_t workers = concurrency;
    if (nthreads > 0 && (uint32_t)nthreads < workers) workers = (uint32_t)nthreads;
    if (workers > n) workers = n ? n : 1;
    std::atomic<uint32_t> next{0};
    std::atomic<int> first_err{DM_OK};
    const uint8_t *base = static_cast<const uint8_t *>(host_base);

    auto worker = [&]() {
        std::vector<uint8_t> buf(zero_copy ? 0 : chunk);       // goproxy's copy buffer
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            const uint64_t len = offsets[i + 1] - offsets[i];
            MemUpstream up(base + offsets[i], len);
            dm::BodyTee tee(e, &up, expect ? expect + 32ull * i : nullptr, len);
            long got;
            if (zero_copy) {
                const void *view = nullptr;
                while ((got = tee.ReadInPlace(&view, chunk)) > 0) {}
            } else {
                while ((got = tee.Read(buf.data(), chunk)) > 0) {}
            }
            if (got < 0 || tee.status() != DM_OK) {
                int exp = DM_OK;
                first_err.compare_exchange_strong(exp, got < 0 ? (int)got : tee.status());
                continue;
            }
            if (digests_out) memcpy(digests_out + 32ull * i, tee.digest(), 32);
            if (matched_out) matched_out[i] = tee.matched() ? 1 : 0;
        }
    };

    const double t0 = now_s();
    std::vector<std::thread> th;
    th.reserve(workers);
    for (uint32_t t = 0; t < workers; ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
    if (seconds) *seconds = now_s() - t0;
    return first_err.load();
}

extern "C" int dm_proxy_serve(dm_engine *e, const uint8_t *digests, uint32_t n, void *host_base,
                              const uint64_t *offsets, size_t chunk, int nthreads, double *seconds)
{
    if (!e || !digests || !offsets || (!host_base && n)) return DM_EINVAL;
    if (chunk == 0) chunk = 32768;
    uint32_t workers = nthreads > 0 ? (uint32_t)nthreads : 1;
    if (workers > n) workers = n ? n : 1;
    std::atomic<uint32_t> next{0};
    std::atomic<int> first_err{DM_OK};
    uint8_t *base = static_cast<uint8_t *>(host_base);

    auto worker = [&]() {
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            dm::HitReader hr(e, digests + 32ull * i);
            int exp = DM_OK;
            if (!hr.hit()) { first_err.compare_exchange_strong(exp, DM_ENOENT); continue; }
            const uint64_t want = offsets[i + 1] - offsets[i];
            if (hr.size() != want) { first_err.compare_exchange_strong(exp, DM_ERANGE); continue; }
            uint8_t *dst = base + offsets[i];
            uint64_t off = 0;
            long got;
            while (off < want && (got = hr.Read(dst + off, (size_t)std::min<uint64_t>(chunk, want - off))) > 0)
                off += (uint64_t)got;
            if (off != want) first_err.compare_exchange_strong(exp, DM_EIO);
        }
    };

    const double t0 = now_s();
    std::vector<std::thread> th;
    th.reserve(workers);
    for (uint32_t t = 0; t < workers; ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
    if (seconds) *seconds = now_s() - t0;
    return first_err.load();
}
