// TEST INFRASTRUCTURE ONLY — multi-threaded soak of the engine's host logic through the C-ABI, linked
// against the fake synchronous CUDA runtime (fake_cuda.cc) so it runs on a CPU-only box under
// ThreadSanitizer / AddressSanitizer.  What it checks: routing of bytes (sequential, ranges, zero-copy,
// unknown size), padding contract, verdicts, cache reads, followers, eviction under a tiny arena, ring
// back-pressure with a tiny ring, checkpoint/resume, verify-only streams, the disk tier, leak
// accounting — with real threads, so the sanitizers see the engine's locking as it really runs.
// Digest arithmetic itself is NOT what this rig proves (the fake kernels use the oracle); the GPU
// parity tests do that.
#include "../../include/demodel_b200.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <random>
#include <string>
#include <thread>
#include <vector>

extern "C" void dmo_sha256(const void *data, size_t len, uint8_t out[32]);

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #c, dm_last_error()); \
                                   failures.fetch_add(1); return; } } while (0)

static std::atomic<int> failures{0};
static std::atomic<int> where[64];             // op each worker is in (watchdog report)
static std::atomic<long> ops{0}, enomem{0}, followed{0};
static uint64_t g_arena = 0;
static bool g_disk = false;                      // the engine has a disk tier (cas_dir): suspend / resume_saved are available                   // a body larger than the arena is verified only: no out-of-order ranges

struct Body { std::vector<uint8_t> bytes; uint8_t digest[32]; };
static std::vector<Body> bodies;

static void make_bodies()
{
    const size_t sizes[] = {0, 1, 63, 64, 65, 4096, 65536, 65537, 200000, 300001, 700000, 1500000};
    for (int rep = 0; rep < 2; ++rep)
        for (size_t s : sizes) {
            Body b;
            b.bytes.resize(s);
            dm_synth_fill_host(0xDE40DE1, 500 + bodies.size(), 0, b.bytes.data(), s);
            dmo_sha256(b.bytes.data(), s, b.digest);
            bodies.push_back(std::move(b));
        }
}

static bool g_inject = false;                   // FAKE_CUDA_FAIL_PPM set: copies and launches fail at random
static std::atomic<long> cuda_failed{0};
static bool tolerate(int rc)
{
    if (rc == DM_ENOMEM) { enomem++; return true; }
    if (rc == DM_ECUDA && g_inject) { cuda_failed++; return true; }      // reported, never silently wrong
    return false;
}

// finish (optionally flushing first); a full arena at the last slab is legal: abort and report "skipped"
static int finish_or_skip(dm_engine *e, uint64_t id, bool flush_first, uint8_t got[32], int *matched)
{
    int rc = flush_first ? dm_stream_flush(e, id) : DM_OK;
    if (rc == DM_OK) rc = dm_stream_finish(e, id, got, matched);
    if (rc == DM_ENOMEM) { enomem++; dm_stream_abort(e, id); return 1; }
    if (rc == DM_ECUDA && g_inject) {            // a device fault during this stream: verdict withheld, nothing cached
        cuda_failed++;
        dm_stream_abort(e, id);                  // no-op if finish got far enough to release the id (ids are never reused)
        if (matched && *matched == 1) return DM_EINVAL;                  // ... and it must not have claimed a match
        return 1;
    }
    return rc;
}

static void worker(dm_engine *e, int tid, double seconds, bool verify_only)
{
    std::mt19937_64 rng(1000 + tid);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(seconds);
    std::vector<uint8_t> scratch(2 << 20);
    while (std::chrono::steady_clock::now() < deadline && !failures.load()) {
        const Body &b = bodies[rng() % bodies.size()];
        const size_t n = b.bytes.size();
        const uint8_t *p = b.bytes.data();
        const bool vo = verify_only || n > g_arena;                  // larger than the arena: verified, never stored
        uint8_t got[32];
        int matched = -1, rc;
        uint64_t id = 0;
        const int op = (int)(rng() % 21);
        ops++;
        where[tid & 63] = op * 1000 + (int)(n >> 10);
        if (op <= 1) {                                               // sequential, random piece size
            rc = dm_stream_open(e, b.digest, (rng() & 1) ? n : 0, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            size_t pos = 0;
            bool dead = false;
            while (pos < n) {
                const size_t k = std::min<size_t>(n - pos, 1 + rng() % 90000);
                rc = dm_stream_write(e, id, p + pos, k);
                if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); dead = true; break; }
                pos += k;
            }
            if (dead) continue;
            rc = finish_or_skip(e, id, rng() & 1, got, &matched);
            if (rc == 1) continue;
            CHECK(rc == DM_OK);
            CHECK(matched == 1 && memcmp(got, b.digest, 32) == 0);
        } else if (op == 2 && n > 1000 && !vo) {            // three range parts, pieces interleaved
            rc = dm_stream_open(e, b.digest, n, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            size_t cut1 = 1 + rng() % (n - 2), cut2 = cut1 + rng() % (n - cut1);
            size_t lo[3] = {0, cut1, cut2}, hi[3] = {cut1, cut2, n};
            bool dead = false;
            while (!dead && (lo[0] < hi[0] || lo[1] < hi[1] || lo[2] < hi[2])) {
                const int k = (int)(rng() % 3);
                if (lo[k] >= hi[k]) continue;
                const size_t m = std::min<size_t>(hi[k] - lo[k], 1 + rng() % 70000);
                rc = dm_stream_write_at(e, id, lo[k], p + lo[k], m);
                if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); dead = true; }
                lo[k] += m;
            }
            if (dead) continue;
            rc = finish_or_skip(e, id, false, got, &matched);
            if (rc == 1) continue;
            CHECK(rc == DM_OK);
            CHECK(matched == 1 && memcmp(got, b.digest, 32) == 0);
        } else if (op == 3) {                                        // zero-copy windows
            rc = dm_stream_open(e, b.digest, n, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            size_t pos = 0;
            bool dead = false;
            while (pos < n) {
                void *win = nullptr;
                size_t cap = 0;
                rc = dm_stream_acquire(e, id, &win, &cap);
                if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); dead = true; break; }
                const size_t k = std::min<size_t>(std::min(cap, n - pos), 1 + rng() % 50000);
                memcpy(win, p + pos, k);
                rc = dm_stream_commit(e, id, k);
                if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); dead = true; break; }
                pos += k;
            }
            if (dead) continue;
            rc = finish_or_skip(e, id, false, got, &matched);
            if (rc == 1) continue;
            CHECK(rc == DM_OK);
            CHECK(matched == 1 && memcmp(got, b.digest, 32) == 0);
        } else if (op == 4) {                                        // client goes away
            rc = dm_stream_open(e, b.digest, n, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            rc = dm_stream_write(e, id, p, n / 2);
            CHECK(rc == DM_OK || tolerate(rc));
            CHECK(dm_stream_abort(e, id) == DM_OK);
        } else if (op == 5) {                                        // wrong oid must be caught
            uint8_t wrong[32];
            memset(wrong, 0x5a, 32);
            rc = dm_stream_open(e, wrong, n, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            rc = dm_stream_write(e, id, p, n);
            if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            rc = finish_or_skip(e, id, false, got, &matched);
            if (rc == 1) continue;
            CHECK(rc == DM_OK);
            CHECK(matched == 0 && memcmp(got, b.digest, 32) == 0);
        } else if (op == 6 || op == 7) {                             // hit serving, io.Copy style + random reads
            uint64_t rid = 0, size = 0;
            if (dm_cache_open(e, b.digest, &rid, &size) != DM_OK) continue;
            CHECK(size == n);
            size_t off = 0, nread = 0;
            const size_t piece = (op == 6) ? 32768 : 1 + rng() % 300000;
            while (off < n) {
                rc = dm_cache_read(e, rid, off, scratch.data(), std::min(piece, scratch.size()), &nread);
                if (rc != DM_OK) { CHECK(tolerate(rc)); break; }                      // injected D2H fault: reported, reader still closes
                CHECK(nread > 0 && memcmp(scratch.data(), p + off, nread) == 0);
                off += nread;
                if (op == 7 && n) off = std::min<size_t>(n, off + rng() % 1000);     // skip around a little
            }
            char meta[512];
            size_t mlen = 0;
            CHECK(dm_cache_meta(e, rid, meta, sizeof meta, &mlen) == DM_OK && mlen > 10);
            CHECK(dm_cache_close(e, rid) == DM_OK);
        } else if (op == 8 && !vo) {                        // coalesce onto an in-flight body
            uint64_t rid = 0, hint = 0;
            if (dm_cache_follow(e, b.digest, &rid, &hint) != DM_OK) continue;
            size_t off = 0, nread = 0;
            bool ok = true;
            for (;;) {
                rc = dm_cache_read(e, rid, off, scratch.data(), 100000, &nread);
                if (rc != DM_OK) { CHECK(rc == DM_ESTATE || rc == DM_ENOENT || tolerate(rc)); ok = false; break; }
                if (nread == 0) break;
                CHECK(off + nread <= n && memcmp(scratch.data(), p + off, nread) == 0);
                off += nread;
            }
            if (ok) { CHECK(off == n); followed++; }
            CHECK(dm_cache_close(e, rid) == DM_OK);
        } else if (op == 9) {
            rc = dm_cache_evict(e, b.digest);
            CHECK(rc == DM_OK || rc == DM_ENOENT || rc == DM_ESTATE);
        } else if (op == 10 && n > 70000) {                          // interrupted download, resumed
            const size_t cut = 65000 + rng() % (n - 65000);
            rc = dm_stream_open(e, b.digest, n, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            rc = dm_stream_write(e, id, p, cut);
            if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            dm_checkpoint ck;
            rc = dm_stream_checkpoint(e, id, &ck);
            if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            CHECK(ck.bytes <= cut && ck.bytes % 64 == 0 && (vo || ck.bytes == cut / 64 * 64));
            CHECK(dm_stream_abort(e, id) == DM_OK);
            rc = dm_stream_resume(e, &ck, b.digest, n, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            if ((rng() & 1) && !vo && ck.bytes) {
                rc = dm_stream_write_at(e, id, 0, p, ck.bytes);
                if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            }
            rc = dm_stream_write(e, id, p + ck.bytes, n - ck.bytes);
            if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            rc = finish_or_skip(e, id, false, got, &matched);
            if (rc == 1) continue;
            CHECK(rc == DM_OK);
            CHECK(matched == 1 && memcmp(got, b.digest, 32) == 0);
        } else if (op == 12 && !vo) {                       // a ragged batch already "in HBM": one launch
            // half of the batches come from a small set that is ingested again and again (REPLACE then finds the previous
            // copies cached and rewrites their extents in place), the others are new every time
            const bool again = rng() & 1;
            const uint64_t first = again ? 9000 + 64 * (rng() % 12) : 20000 + rng() % 100000;
            std::mt19937_64 lrng(again ? first : rng());
            const uint32_t nb = 1 + (uint32_t)(lrng() % 24);
            std::vector<uint64_t> off(nb), len(nb);
            uint64_t pos = 0;
            for (uint32_t i = 0; i < nb; ++i) {
                off[i] = pos;
                len[i] = (lrng() % 6 == 0) ? 0 : 16 * (lrng() % 3000) + (lrng() % 3 == 0 ? lrng() % 16 : 0);
                pos += (len[i] + 15) / 16 * 16;
            }
            std::vector<uint8_t> dev(pos + 16);                      // the rig's device memory is host memory
            rc = dm_synth_fill_device_many(e, 0xDE40DE1, first, dev.data(), off.data(), len.data(), nb);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            std::vector<uint8_t> want(32 * nb), digs(32 * nb), mat(nb, 7);
            for (uint32_t i = 0; i < nb; ++i) dmo_sha256(dev.data() + off[i], len[i], &want[32 * i]);
            {   // the device generator must equal the host generator
                const uint32_t i = (uint32_t)(rng() % nb);
                std::vector<uint8_t> host(len[i]);
                dm_synth_fill_host(0xDE40DE1, first + i, 0, host.data(), len[i]);
                CHECK(len[i] == 0 || memcmp(host.data(), dev.data() + off[i], len[i]) == 0);
            }
            const bool cache = rng() & 1;
            std::vector<uint8_t> expect = want;
            const uint32_t ci = (uint32_t)(rng() % nb);
            const bool corrupt = rng() % 4 == 0;
            if (corrupt) expect[32 * ci] ^= 1;
            uint32_t flags = cache ? DM_ING_REPLACE : DM_ING_HASH_ONLY;
            const uint32_t pick = (uint32_t)(rng() % 8);
            if (pick >= 1 && pick <= 6) flags |= pick << DM_ING_SPW_SHIFT;      // 1, 2, 4, 8, 16, 32 streams per warp
            double ms = -1;
            rc = dm_ingest_device(e, dev.data(), off.data(), len.data(), nb, expect.data(), digs.data(), mat.data(), flags, &ms);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            CHECK(memcmp(digs.data(), want.data(), 32 * nb) == 0);
            for (uint32_t i = 0; i < nb; ++i) {
                const bool same_as_ci = memcmp(&want[32 * i], &want[32 * ci], 32) == 0;   // duplicates (empty blobs) share the verdict
                if (!(corrupt && same_as_ci)) CHECK(mat[i] == 1);
                if (corrupt && i == ci) CHECK(mat[i] == 0);
            }
            if (cache) {
                const uint32_t i = (uint32_t)(rng() % nb);
                uint64_t rid = 0, size = 0;
                if (!(corrupt && i == ci) && dm_cache_open(e, &want[32 * i], &rid, &size) == DM_OK) {
                    CHECK(size == len[i]);
                    void *ptrs[16];
                    uint64_t lens[16];
                    const int ne = dm_cache_device_extents(e, rid, ptrs, lens, 16);
                    CHECK((ne >= 0 && ne <= 16) || ne == DM_ESTATE);       // DM_ESTATE: already evicted to the disk tier
                    uint64_t seen = ne < 0 ? size : 0;
                    for (int x = 0; x < ne; ++x) {                   // in the rig a device pointer can be read directly
                        CHECK(memcmp(ptrs[x], dev.data() + off[i] + seen, lens[x]) == 0);
                        seen += lens[x];
                    }
                    CHECK(seen == size);
                    CHECK(dm_cache_close(e, rid) == DM_OK);
                }
            }
        } else if (op == 13) {                                       // response headers travel with the blob
            rc = dm_stream_open(e, b.digest, n, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            CHECK(dm_stream_set_meta(e, id, "Content-Type", "application/octet-stream") == DM_OK);
            CHECK(dm_stream_set_meta(e, id, "ETag", "\"quoted\\value\"") == DM_OK);
            rc = dm_stream_write(e, id, p, n);
            if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            rc = finish_or_skip(e, id, false, got, &matched);
            if (rc == 1) continue;
            CHECK(rc == DM_OK && matched == 1);
            uint64_t rid = 0, size = 0;
            if (!vo && dm_cache_open(e, b.digest, &rid, &size) == DM_OK) {
                char meta[1024];
                size_t mlen = 0;
                CHECK(dm_cache_meta(e, rid, meta, sizeof meta, &mlen) == DM_OK && mlen < sizeof meta);
                CHECK(strstr(meta, "\"digest\":\"sha256:") != nullptr);
                CHECK(dm_cache_close(e, rid) == DM_OK);
            }
        } else if (op == 14) {                                       // API abuse: errors, never a crash or a leak
            const uint64_t bogus = 0x7000000000ull + rng() % 1000;
            size_t nread = 0, cap = 0;
            void *win = nullptr;
            CHECK(dm_stream_write(e, bogus, scratch.data(), 1) < 0);
            CHECK(dm_stream_write_at(e, bogus, 0, scratch.data(), 1) < 0);
            CHECK(dm_stream_finish(e, bogus, got, &matched) < 0);
            CHECK(dm_stream_flush(e, bogus) < 0);
            CHECK(dm_stream_abort(e, bogus) < 0);
            CHECK(dm_stream_acquire(e, bogus, &win, &cap) < 0);
            CHECK(dm_stream_commit(e, bogus, 1) < 0);
            CHECK(dm_stream_set_meta(e, bogus, "k", "v") < 0);
            CHECK(dm_cache_read(e, bogus, 0, scratch.data(), 10, &nread) < 0);
            CHECK(dm_cache_close(e, bogus) < 0);
            CHECK(dm_stream_open(nullptr, b.digest, n, &id) < 0);
            CHECK(dm_stream_open(e, b.digest, n, nullptr) < 0);
            CHECK(dm_cache_open(e, nullptr, &id, nullptr) < 0);
            dm_checkpoint junk;
            memset(&junk, 0xee, sizeof junk);
            CHECK(dm_stream_resume(e, &junk, b.digest, n, &id) < 0);
            rc = dm_stream_open(e, nullptr, 0, &id);                 // unknown size, no expectation
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            rc = dm_stream_acquire(e, id, &win, &cap);
            if (rc == DM_OK) {
                CHECK(cap > 0);
                CHECK(dm_stream_acquire(e, id, &win, &cap) == DM_ESTATE);       // one window at a time
                CHECK(dm_stream_write(e, id, scratch.data(), 1) == DM_ESTATE);
                CHECK(dm_stream_commit(e, id, cap + 1) == DM_EINVAL);
                CHECK(dm_stream_commit(e, id, 0) == DM_OK);
                CHECK(dm_stream_commit(e, id, 0) == DM_ESTATE);
                CHECK(dm_stream_write_at(e, id, ~0ull - 10, scratch.data(), 100) == DM_ERANGE);
            } else {
                CHECK(tolerate(rc));
            }
            if (n) { rc = dm_stream_write(e, id, p, std::min<size_t>(n, 5000)); CHECK(rc == DM_OK || tolerate(rc)); }
            if (rc != DM_OK) { dm_stream_abort(e, id); continue; }
            rc = finish_or_skip(e, id, false, got, &matched);
            if (rc == 1) continue;
            CHECK(rc == DM_OK && matched == 1);                      // no expectation: whatever arrived is accepted
            CHECK(dm_stream_finish(e, id, got, &matched) < 0);       // finish released the id
            CHECK(dm_stream_write(e, id, scratch.data(), 1) < 0);
            CHECK(dm_stream_abort(e, id) < 0);
            uint64_t rid = 0, size = 0;
            if (!vo && dm_cache_open(e, got, &rid, &size) == DM_OK) {
                CHECK(size == std::min<size_t>(n, 5000));
                CHECK(dm_cache_read(e, rid, size + 1, scratch.data(), 1, &nread) == DM_ERANGE);
                CHECK(dm_cache_read(e, rid, size, scratch.data(), 1, &nread) == DM_OK && nread == 0);      // EOF
                rc = dm_cache_evict(e, got);
                CHECK(rc == DM_ESTATE || rc == DM_ENOENT || rc == DM_OK);        // busy here; another thread may hold or have evicted it
                CHECK(dm_cache_close(e, rid) == DM_OK);
                CHECK(dm_cache_close(e, rid) < 0);
                CHECK(dm_cache_read(e, rid, 0, scratch.data(), 1, &nread) < 0);
            }
        } else if (op == 15) {                                       // URL-keyed fetch (digest unknown up front) and hit by URL
            char url[160];
            snprintf(url, sizeof url, "https://huggingface.co/org/model/resolve/main/file-%zu-%d.safetensors", (size_t)(&b - &bodies[0]), (int)(rng() % 3));
            rc = dm_proxy_fetch(e, url, p, n, (rng() & 1) ? b.digest : nullptr, 1 + rng() % 60000, got, &matched);
            if (rc == DM_ENOMEM) { enomem++; continue; }
            if (rc == DM_ECUDA && g_inject) { cuda_failed++; continue; }
            CHECK(rc == DM_OK);
            CHECK(matched == 1 && memcmp(got, b.digest, 32) == 0);
            uint8_t ad[32];
            CHECK(dm_cache_alias_get(e, url, ad) == DM_OK && memcmp(ad, b.digest, 32) == 0);
            uint64_t rid = 0, size = 0;
            rc = dm_proxy_request(e, url, &rid, &size);
            if (rc == DM_OK) {                                       // (a miss is legal: verify-only engine, or evicted meanwhile)
                CHECK(size == n);
                size_t nread = 0;
                if (n) {
                    const size_t off = rng() % n;
                    rc = dm_cache_read(e, rid, off, scratch.data(), std::min<size_t>(n - off, 50000), &nread);
                    if (rc == DM_OK) CHECK(nread > 0 && memcmp(scratch.data(), p + off, nread) == 0);
                    else CHECK(tolerate(rc));
                }
                CHECK(dm_cache_close(e, rid) == DM_OK);
            } else CHECK(rc == DM_ENOENT);
            // an OCI URL names its digest: resolved without the index
            char oci[200] = "https://registry.ollama.ai/v2/library/x/blobs/sha256:";
            for (int i = 0; i < 32; ++i) snprintf(oci + strlen(oci), 3, "%02x", b.digest[i]);
            rc = dm_proxy_request(e, oci, &rid, &size);
            if (rc == DM_OK) { CHECK(size == n); CHECK(dm_cache_close(e, rid) == DM_OK); } else CHECK(rc == DM_ENOENT);
            CHECK(dm_cache_alias_get(e, "https://never/seen", ad) == DM_ENOENT);
            CHECK(dm_cache_alias_put(e, "bad\nkey", b.digest) == DM_EINVAL);
        } else if (op == 16) {                                       // interrupted download saved to the disk tier and picked up again
            // a body private to this thread: the saved files are named by the expected digest
            std::vector<uint8_t> mine(70000 + rng() % 400000);
            dm_synth_fill_host(0xDE40DE1, 70000 + (uint64_t)tid * 1000 + rng() % 8, 0, mine.data(), mine.size());
            uint8_t md[32];
            dmo_sha256(mine.data(), mine.size(), md);
            const size_t mn = mine.size(), cut = 1 + rng() % (mn - 1);
            rc = dm_stream_open(e, md, mn, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            rc = dm_stream_write(e, id, mine.data(), cut);
            if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            uint64_t from = ~0ull;
            rc = dm_stream_suspend(e, id, &from);
            if (!g_disk || vo || mn > g_arena) { CHECK(rc == DM_ESTATE); CHECK(dm_stream_abort(e, id) == DM_OK); continue; }
            if (rc != DM_OK) { CHECK(tolerate(rc) || rc == DM_EIO); dm_stream_abort(e, id); continue; }
            CHECK(from <= cut && from % 64 == 0 && from == cut / 64 * 64);
            CHECK(dm_stream_write(e, id, mine.data(), 1) < 0);       // suspend released the id
            uint64_t from2 = ~0ull;
            rc = dm_stream_resume_saved(e, md, mn, &id, &from2);
            if (rc != DM_OK) { CHECK(tolerate(rc)); continue; }       // (arena full while re-loading the prefix)
            CHECK(from2 == from);
            rc = dm_stream_write(e, id, mine.data() + from, mn - from);
            if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            rc = finish_or_skip(e, id, false, got, &matched);
            if (rc == 1) continue;
            CHECK(rc == DM_OK);
            CHECK(matched == 1 && memcmp(got, md, 32) == 0);
            uint64_t rid = 0, size = 0;
            if (dm_cache_open(e, md, &rid, &size) == DM_OK) {        // cached WHOLE: the saved prefix came back from disk
                CHECK(size == mn);
                size_t nread = 0;
                rc = dm_cache_read(e, rid, 0, scratch.data(), std::min<size_t>(mn, scratch.size()), &nread);
                if (rc == DM_OK) CHECK(nread == std::min<size_t>(mn, scratch.size()) && memcmp(scratch.data(), mine.data(), nread) == 0);
                else CHECK(tolerate(rc));
                CHECK(dm_cache_close(e, rid) == DM_OK);
                dm_cache_evict(e, md);
            }
            CHECK(dm_stream_resume_saved(e, md, mn, &id, &from2) == DM_ENOENT);      // the saved files were consumed
        } else if (op == 17) {                                       // error text by id, from another thread (cgo: goroutines migrate)
            const uint64_t bogus = 0x7100000000ull + (uint64_t)tid * 4096 + rng() % 1000;
            CHECK(dm_stream_write(e, bogus, scratch.data(), 1) == DM_EINVAL);
            const std::string here = dm_last_error();
            std::string there;
            std::thread([&] { char t[256]; size_t l = 0; if (dm_error_detail(e, bogus, t, sizeof t, &l) == DM_OK) there.assign(t, std::min(l, sizeof t - 1)); }).join();
            CHECK(!here.empty() && here == there);
            char tiny[4];
            size_t full = 0;
            CHECK(dm_error_detail(e, bogus, tiny, sizeof tiny, &full) == DM_OK && full == here.size() && strlen(tiny) == 3);
        } else if (op == 18 || op == 19) {                           // finish racing abort (18) / a second finish (19) on one id
            rc = dm_stream_open(e, b.digest, n, &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            rc = dm_stream_write(e, id, p, n);
            if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            int rc2 = 0, m2 = -1;
            uint8_t got2[32];
            std::thread other([&] {
                if (rng() & 1) std::this_thread::yield();
                rc2 = op == 18 ? dm_stream_abort(e, id) : dm_stream_finish(e, id, got2, &m2);
            });
            rc = dm_stream_finish(e, id, got, &matched);
            other.join();
            // whoever loses sees a closed / unknown stream; nobody hangs, nothing leaks (checked at the end of the soak)
            if (rc == DM_OK) CHECK(matched == 1 && memcmp(got, b.digest, 32) == 0);
            else { CHECK(rc == DM_ESTATE || rc == DM_EINVAL || tolerate(rc)); if (rc == DM_ENOMEM || rc == DM_ECUDA) dm_stream_abort(e, id); }
            if (op == 19 && rc2 == DM_OK) CHECK(m2 == 1 && memcmp(got2, b.digest, 32) == 0);
            if (rc2 != DM_OK) CHECK(rc2 == DM_ESTATE || rc2 == DM_EINVAL || tolerate(rc2));
        } else if (op == 20 && !vo) {                                // a STREAM carries a blob of the re-ingested sets (op 12):
            // while a batch has that blob parked, publication of this body re-homes the entry and the batch must yield
            const uint64_t first = 9000 + 64 * (rng() % 12);
            std::mt19937_64 lrng(first);
            const uint32_t nb = 1 + (uint32_t)(lrng() % 24);
            std::vector<uint64_t> len(nb);
            for (uint32_t i = 0; i < nb; ++i) len[i] = (lrng() % 6 == 0) ? 0 : 16 * (lrng() % 3000) + (lrng() % 3 == 0 ? lrng() % 16 : 0);
            const uint32_t i = (uint32_t)(rng() % nb);
            std::vector<uint8_t> mine(len[i] + 1);                   // (+1: data() of an empty vector may be null)
            dm_synth_fill_host(0xDE40DE1, first + i, 0, mine.data(), len[i]);
            uint8_t md[32];
            dmo_sha256(mine.data(), len[i], md);
            rc = dm_stream_open(e, md, len[i], &id);
            if (tolerate(rc)) continue;
            CHECK(rc == DM_OK);
            rc = dm_stream_write(e, id, mine.data(), len[i]);
            if (rc != DM_OK) { CHECK(tolerate(rc)); dm_stream_abort(e, id); continue; }
            rc = finish_or_skip(e, id, false, got, &matched);
            if (rc == 1) continue;
            CHECK(rc == DM_OK && matched == 1 && memcmp(got, md, 32) == 0);
            uint64_t rid = 0, size = 0;
            if (dm_cache_open(e, md, &rid, &size) == DM_OK) {        // (it may be parked by a batch right now: a miss is fine)
                CHECK(size == len[i]);
                size_t nread = 0;
                rc = dm_cache_read(e, rid, 0, scratch.data(), std::min<size_t>(len[i], scratch.size()), &nread);
                if (rc == DM_OK) CHECK(nread == std::min<size_t>(len[i], scratch.size()) && memcmp(scratch.data(), mine.data(), nread) == 0);
                else CHECK(tolerate(rc));
                CHECK(dm_cache_close(e, rid) == DM_OK);
            }
        } else if (op == 11) {                                       // metadata + stats are always safe to call
            dm_stats st;
            CHECK(dm_engine_stats(e, &st) == DM_OK);
            CHECK(st.ring_slabs_free <= st.ring_slabs_total);
            uint64_t sz = 0;
            rc = dm_cache_contains(e, b.digest, &sz);
            CHECK(rc == DM_ENOENT || (rc == DM_OK && sz == n));
        }
    }
}

// The goroutine-model driver (BodyTee / HitReader over the C-ABI) in all its modes, then the
// manifest-aware prefetch, on an engine of its own so results are deterministic.
static void driver_phase(bool verify_only)
{
    dm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    const char *drv_ring = getenv("RIG_DRV_RING");              // e.g. 262144: 4 slabs for 8 interleaved bodies per thread
    cfg.hbm_cas_bytes = 64u << 20;
    cfg.ring_bytes = drv_ring && *drv_ring ? strtoull(drv_ring, nullptr, 0) : 2u << 20;
    cfg.slab_bytes = 64u << 10;
    cfg.max_streams = 256;
    cfg.flags = verify_only ? DM_F_NO_HBM_CAS : 0;
    dm_engine *e = nullptr;
    CHECK(dm_engine_create(&cfg, &e) == DM_OK);
    const uint32_t n = (uint32_t)bodies.size();
    std::vector<uint64_t> off(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) off[i + 1] = off[i] + bodies[i].bytes.size();
    std::vector<uint8_t> host(off[n] ? off[n] : 1), expect(32ull * n), digs(32ull * n), verdict(n), back(host.size());
    for (uint32_t i = 0; i < n; ++i) {
        if (!bodies[i].bytes.empty()) memcpy(host.data() + off[i], bodies[i].bytes.data(), bodies[i].bytes.size());
        memcpy(expect.data() + 32ull * i, bodies[i].digest, 32);
    }
    for (int mode = 0; mode <= 4; ++mode) {
        double secs = 0;
        std::fill(verdict.begin(), verdict.end(), 9);
        const int rc = dm_proxy_drive(e, host.data(), off.data(), n, expect.data(), 4096 + 1000 * mode, 8, mode < 2 ? 3 : 0, mode,
                                      digs.data(), verdict.data(), &secs);
        CHECK(rc == DM_OK);
        for (uint32_t i = 0; i < n; ++i) {
            if (mode == 4) { CHECK(verdict[i] == 2); continue; }
            CHECK(verdict[i] == 1 && memcmp(digs.data() + 32ull * i, bodies[i].digest, 32) == 0);
        }
        if (mode != 4 && !verify_only) {
            std::fill(back.begin(), back.end(), 0);
            CHECK(dm_proxy_serve(e, expect.data(), n, back.data(), off.data(), 10000, 4, &secs) == DM_OK);
            CHECK(memcmp(back.data(), host.data(), off[n]) == 0);
        }
        for (uint32_t i = 0; i < n; ++i) dm_cache_evict(e, bodies[i].digest);
    }
    if (!verify_only) {                                     // manifest: parse, prefetch, pull, second prefetch = all hits
        auto hex = [](const uint8_t *d) { std::string s; char t[3]; for (int i = 0; i < 32; ++i) { snprintf(t, 3, "%02x", d[i]); s += t; } return s; };
        std::string m = "{\"schemaVersion\":2,\"config\":{\"digest\":\"sha256:" + hex(bodies[5].digest) + "\",\"size\":" +
                        std::to_string(bodies[5].bytes.size()) + "},\"layers\":[";
        for (int k = 6; k < 10; ++k)
            m += std::string(k > 6 ? "," : "") + "{\"mediaType\":\"l\",\"digest\":\"sha256:" + hex(bodies[k].digest) + "\",\"size\":" +
                 std::to_string(bodies[k].bytes.size()) + "}";
        m += "]}";
        dm_layer L[8];
        uint32_t nl = 0;
        CHECK(dm_manifest_parse(m.data(), m.size(), L, 8, &nl) == DM_OK && nl == 5);
        uint64_t ids[8];
        CHECK(dm_manifest_prefetch(e, L, nl, ids) == DM_OK);
        for (uint32_t i = 0; i < nl; ++i) {
            CHECK(ids[i] != 0);
            const Body &b = bodies[5 + i];
            CHECK(dm_stream_write(e, ids[i], b.bytes.data(), b.bytes.size()) == DM_OK);
            uint8_t got[32];
            int matched = 0;
            CHECK(dm_stream_finish(e, ids[i], got, &matched) == DM_OK && matched == 1);
        }
        CHECK(dm_manifest_prefetch(e, L, nl, ids) == DM_OK);
        for (uint32_t i = 0; i < nl; ++i) CHECK(ids[i] == 0);
    }
    dm_stats st;
    dm_engine_stats(e, &st);
    CHECK(st.open_streams == 0 && st.open_readers == 0);
    dm_engine_destroy(e);
}

// a partial slab recalled by the PUMP whose DMA cannot be placed (arena full while an
// unknown-size body grows) used to drop the bytes silently - the writer carried on and finish() succeeded over a
// truncated body.  Built deterministically: arena of exactly three slabs, ring of four.
static void lost_bytes_are_sticky()
{
    dm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    const uint32_t slab = 64u << 10;
    cfg.hbm_cas_bytes = 3 * slab; cfg.ring_bytes = 4 * slab; cfg.slab_bytes = slab; cfg.max_streams = 16;
    dm_engine *e = nullptr;
    CHECK(dm_engine_create(&cfg, &e) == DM_OK);
    std::vector<uint8_t> buf(4 * slab, 0x5c);
    uint64_t a = 0, b = 0, c = 0;
    CHECK(dm_stream_open(e, nullptr, 0, &a) == DM_OK);                       // A: size unknown
    CHECK(dm_stream_write(e, a, buf.data(), slab + 100) == DM_OK);           // one slab sent (extent: 1 slab), 100 B staged
    CHECK(dm_stream_open(e, nullptr, 2 * slab, &b) == DM_OK);                // B: reserves the rest of the arena
    CHECK(dm_stream_write_at(e, b, 1000, buf.data(), 10) == DM_OK);          // ... and pins the other three ring slabs
    CHECK(dm_stream_write_at(e, b, 5000, buf.data(), 10) == DM_OK);
    CHECK(dm_stream_write_at(e, b, 9000, buf.data(), 10) == DM_OK);
    CHECK(dm_stream_open(e, nullptr, 0, &c) == DM_OK);
    int rc_c = 1;
    std::thread t([&] { rc_c = dm_stream_write(e, c, buf.data(), 10); });    // blocks: the ring is empty and nothing is in flight
    int rc = DM_OK;
    for (int i = 0; i < 4000 && rc == DM_OK; ++i) {                          // the pump recalls A's 100 bytes: no room for them
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
        rc = dm_stream_write(e, a, buf.data(), 0);
    }
    t.join();
    CHECK(rc == DM_ENOMEM);                                                  // the writer is told on its next call ...
    uint8_t got[32];
    int matched = -1;
    CHECK(dm_stream_finish(e, a, got, &matched) == DM_ENOMEM);               // ... and the truncated body can not be finished
    char why[160];
    size_t n = 0;
    CHECK(dm_error_detail(e, a, why, sizeof why, &n) == DM_OK && strstr(why, "dropped") != nullptr);
    CHECK(dm_stream_abort(e, a) == DM_OK);
    CHECK(rc_c == DM_OK);                                                    // C got a recalled slab
    CHECK(dm_stream_abort(e, c) == DM_OK);
    CHECK(dm_stream_abort(e, b) == DM_OK);
    dm_stats st;
    for (int i = 0; i < 500; ++i) { dm_engine_stats(e, &st); if (st.ring_slabs_free == st.ring_slabs_total) break; std::this_thread::sleep_for(std::chrono::milliseconds(2)); }
    CHECK(st.ring_slabs_free == st.ring_slabs_total && st.open_streams == 0 && st.blobs_committed == 0);
    dm_engine_destroy(e);
}

// dm_ingest_device over blobs that are already cached (DM_ING_REPLACE): the previous copies are parked and their extents
// written again in place, in chunks on streams of their own when the batch is large (forced here with
// DM_INGEST_CHUNKS).  The arena has room for ONE copy of the batch plus a little: evicting and re-allocating out of
// order, or leaking a parked extent, shows as DM_ENOMEM or as a changed hbm_cas_used.
static void bulk_replace_rewrites_cached_blobs_in_place()
{
    setenv("DM_INGEST_CHUNKS", "3", 1);
    const uint32_t n = 1500;
    std::vector<uint64_t> off(n), len(n);
    std::mt19937_64 rng(77);
    uint64_t pos = 0, arena = 0;
    for (uint32_t i = 0; i < n; ++i) {
        off[i] = pos;
        len[i] = i % 97 == 0 ? 0 : 16 * (rng() % 200) + (i % 5 == 0 ? rng() % 16 : 0);      // ragged, some empty
        pos += (len[i] + 15) / 16 * 16;
        arena += (std::max<uint64_t>(len[i], 1) + 255) / 256 * 256;
    }
    dm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.hbm_cas_bytes = arena + (64u << 10); cfg.ring_bytes = 1u << 20; cfg.slab_bytes = 64u << 10; cfg.max_streams = 64;
    dm_engine *e = nullptr;
    CHECK(dm_engine_create(&cfg, &e) == DM_OK);
    unsetenv("DM_INGEST_CHUNKS");
    std::vector<uint8_t> dev(pos + 16);
    CHECK(dm_synth_fill_device_many(e, 0xDE40DE1, 70000, dev.data(), off.data(), len.data(), n) == DM_OK);
    std::vector<uint8_t> want(32 * n), digs(32 * n), mat(n);
    for (uint32_t i = 0; i < n; ++i) dmo_sha256(dev.data() + off[i], len[i], &want[32 * i]);
    auto check_cached = [&](uint32_t i, bool present) {
        uint64_t rid = 0, size = 0;
        const int rc = dm_cache_open(e, &want[32 * i], &rid, &size);
        if (!present) { CHECK(rc == DM_ENOENT); return; }
        CHECK(rc == DM_OK && size == len[i]);
        void *ptrs[4];
        uint64_t lens[4];
        const int ne = dm_cache_device_extents(e, rid, ptrs, lens, 4);
        CHECK(ne == 1 && lens[0] == len[i] && (len[i] == 0 || memcmp(ptrs[0], dev.data() + off[i], len[i]) == 0));
        CHECK(dm_cache_close(e, rid) == DM_OK);
    };
    dm_stats st0, st;
    for (int round = 0; round < 4; ++round) {
        // round 0: everything is new; 1: lane per stream, chunked; 2: 8 streams per warp (one launch, jobs sorted by
        // length, so positions differ from the caller's indices); 3: the kernel the count picks
        const uint32_t shape = round == 1 ? 6u << DM_ING_SPW_SHIFT : round == 2 ? 4u << DM_ING_SPW_SHIFT : 0;
        uint64_t pinned = 0, psize = 0;
        if (round == 2) CHECK(dm_cache_open(e, &want[32 * 11], &pinned, &psize) == DM_OK);      // a blob with a reader stays where it is
        std::fill(mat.begin(), mat.end(), 7);
        CHECK(dm_ingest_device(e, dev.data(), off.data(), len.data(), n, want.data(), digs.data(), mat.data(), DM_ING_REPLACE | shape, nullptr) == DM_OK);
        CHECK(memcmp(digs.data(), want.data(), 32 * n) == 0);
        for (uint32_t i = 0; i < n; ++i) CHECK(mat[i] == 1);
        if (pinned) CHECK(dm_cache_close(e, pinned) == DM_OK);
        CHECK(dm_engine_stats(e, &st) == DM_OK);
        if (round == 0) st0 = st;
        CHECK(st.hbm_cas_used == st0.hbm_cas_used);
        if (round == 1) CHECK(st.kernel_launches - st0.kernel_launches == 3 && st.launches_wide == 3);
        for (uint32_t i = 0; i < n; i += 37) check_cached(i, true);
        check_cached(11, true);
        check_cached(n - 1, true);
    }
    // the bytes of one blob change under its cached copy: the re-ingest reports the mismatch and that blob is gone
    const uint32_t m = 501;
    CHECK(len[m] > 0);
    dev[off[m]] ^= 0x40;
    CHECK(dm_ingest_device(e, dev.data(), off.data(), len.data(), n, want.data(), digs.data(), mat.data(), DM_ING_REPLACE | (6u << DM_ING_SPW_SHIFT), nullptr) == DM_OK);
    for (uint32_t i = 0; i < n; ++i) CHECK(mat[i] == (i == m ? 0 : 1));
    check_cached(m, false);
    check_cached(m - 1, true);
    check_cached(m + 1, true);
    CHECK(dm_engine_stats(e, &st) == DM_OK);
    CHECK(st.hbm_cas_used == st0.hbm_cas_used - (len[m] + 255) / 256 * 256 && st.blobs_mismatched == 1);
    dmo_sha256(dev.data() + off[m], len[m], &want[32 * m]);                  // what is there now goes in as a new blob
    CHECK(dm_ingest_device(e, dev.data(), off.data(), len.data(), n, want.data(), digs.data(), mat.data(), DM_ING_REPLACE, nullptr) == DM_OK);
    for (uint32_t i = 0; i < n; ++i) CHECK(mat[i] == 1);
    check_cached(m, true);
    CHECK(dm_engine_stats(e, &st) == DM_OK);
    CHECK(st.hbm_cas_used == st0.hbm_cas_used && st.open_readers == 0);
    dm_engine_destroy(e);
}

// Shutdown with transfers in flight: the proxy is stopped while bodies are half way and hits are being served.
// Destroy must not hang, crash or touch freed memory (ASan), whatever state the streams and readers are in.
static void destroy_with_open_handles(const char *cas_dir)
{
    dm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.hbm_cas_bytes = 16u << 20;
    cfg.ring_bytes = 512u << 10;
    cfg.slab_bytes = 64u << 10;
    cfg.max_streams = 64;
    cfg.cas_dir = cas_dir;
    dm_engine *e = nullptr;
    CHECK(dm_engine_create(&cfg, &e) == DM_OK);
    const Body &big = bodies.back();
    uint64_t id = 0, rid = 0, size = 0;
    uint8_t got[32];
    int matched = 0;
    CHECK(dm_stream_open(e, big.digest, big.bytes.size(), &id) == DM_OK);                 // a finished, cached blob with an open reader
    CHECK(dm_stream_write(e, id, big.bytes.data(), big.bytes.size()) == DM_OK);
    CHECK(dm_stream_finish(e, id, got, &matched) == DM_OK && matched == 1);
    CHECK(dm_cache_open(e, big.digest, &rid, &size) == DM_OK);
    std::vector<uint8_t> buf(100000);
    size_t nread = 0;
    CHECK(dm_cache_read(e, rid, 0, buf.data(), buf.size(), &nread) == DM_OK);            // read-ahead windows in flight
    const Body &b1 = bodies[bodies.size() - 2], &b2 = bodies[bodies.size() - 3];
    uint64_t s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    CHECK(dm_stream_open(e, b1.digest, b1.bytes.size(), &s1) == DM_OK);                  // half-written body
    CHECK(dm_stream_write(e, s1, b1.bytes.data(), b1.bytes.size() / 2 + 13) == DM_OK);
    CHECK(dm_stream_open(e, b2.digest, b2.bytes.size(), &s2) == DM_OK);                  // range parts with holes
    CHECK(dm_stream_write_at(e, s2, 100000, b2.bytes.data() + 100000, 5000) == DM_OK);
    CHECK(dm_stream_write_at(e, s2, 200000, b2.bytes.data() + 200000, 70001) == DM_OK);
    CHECK(dm_stream_open(e, nullptr, 0, &s3) == DM_OK);                                   // a lent-out zero-copy window
    void *win = nullptr;
    size_t cap = 0;
    CHECK(dm_stream_acquire(e, s3, &win, &cap) == DM_OK);
    CHECK(dm_stream_open(e, b1.digest, 0, &s4) == DM_OK);                                 // flushed, verdict never collected
    CHECK(dm_stream_write(e, s4, b1.bytes.data(), b1.bytes.size()) == DM_OK);
    CHECK(dm_stream_flush(e, s4) == DM_OK);
    uint64_t fr = 0, hint = 0;
    (void)dm_cache_follow(e, b1.digest, &fr, &hint);                                      // a follower attached to s1 (may be refused)
    dm_engine_destroy(e);
}

// `engine_soak create`: one dm_engine_create / destroy.  With FAKE_CUDA_FAIL_ALLOC_NTH=k the k-th allocation fails:
// exit 2 = create failed cleanly, 0 = it succeeded (k is past the last allocation).  Run under ASan + LSan.
static int create_only()
{
    dm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.hbm_cas_bytes = 8u << 20; cfg.ring_bytes = 1u << 20; cfg.slab_bytes = 64u << 10; cfg.max_streams = 64;
    cfg.cas_dir = "/tmp/dm_rig_create_cas";
    dm_engine *e = nullptr;
    const int rc = dm_engine_create(&cfg, &e);
    if (rc != DM_OK) { if (e) { fprintf(stderr, "create failed but returned a handle\n"); return 1; } return 2; }
    dm_engine_destroy(e);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 1 && strcmp(argv[1], "create") == 0) return create_only();
    const double seconds = argc > 1 ? atof(argv[1]) : 5.0;
    const int threads = argc > 2 ? atoi(argv[2]) : 6;
    const char *cas_dir = argc > 3 && argv[3][0] ? argv[3] : nullptr;
    const bool verify_only = argc > 4 && atoi(argv[4]) != 0;
    make_bodies();
    g_inject = getenv("FAKE_CUDA_FAIL_PPM") != nullptr;
    dm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.device = 0;
    auto env_u64 = [](const char *name, uint64_t dflt) { const char *v = getenv(name); return v && *v ? strtoull(v, nullptr, 0) : dflt; };
    cfg.hbm_cas_bytes = g_arena = env_u64("RIG_ARENA", 6u << 20);   // tiny arena: eviction and ENOMEM are part of the test
    cfg.ring_bytes = env_u64("RIG_RING", 1u << 20);          // 16 slabs of 64 KiB for up to `threads` writers: back-pressure
    cfg.slab_bytes = (uint32_t)env_u64("RIG_SLAB", 64u << 10);
    cfg.max_streams = 256;
    cfg.cas_dir = cas_dir;
    g_disk = cas_dir != nullptr;
    cfg.flags = (verify_only ? DM_F_NO_HBM_CAS : 0) | DM_F_NUMA_LOCAL;          // (no PCI topology on the rig: the flag is a no-op)
    dm_engine *e = nullptr;
    if (dm_engine_create(&cfg, &e) != DM_OK) { fprintf(stderr, "create failed: %s\n", dm_last_error()); return 1; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(worker, e, t, seconds, verify_only);
    std::atomic<bool> done{false};
    std::thread dog([&] {                        // a stuck engine must fail the test with a report, not hang it
        long last = -1;
        int still = 0;
        while (!done.load()) {
            std::this_thread::sleep_for(std::chrono::milliseconds(500));
            const long now = ops.load();
            still = now == last ? still + 1 : 0;
            last = now;
            if (still == 40) {
                dm_stats w;
                dm_engine_stats(e, &w);
                fprintf(stderr, "STUCK for 20 s: open_streams=%llu open_readers=%llu slabs free/total=%llu/%llu hbm_used=%llu ring_waits=%llu launches=%llu\n",
                        (unsigned long long)w.open_streams, (unsigned long long)w.open_readers, (unsigned long long)w.ring_slabs_free,
                        (unsigned long long)w.ring_slabs_total, (unsigned long long)w.hbm_cas_used, (unsigned long long)w.ring_waits,
                        (unsigned long long)w.kernel_launches);
                for (int t = 0; t < threads && t < 64; ++t) fprintf(stderr, "  worker %d: op %d, body %d KiB\n", t, where[t].load() / 1000, where[t].load() % 1000);
                fprintf(stderr, "ENGINE SOAK FAILED (stuck)\n");
                _Exit(3);
            }
        }
    });
    for (auto &t : th) t.join();
    done = true;
    dog.join();
    dm_stats st;
    for (int i = 0; i < 500; ++i) {
        dm_engine_stats(e, &st);
        if (st.ring_slabs_free == st.ring_slabs_total && st.free_stream_slots == 256) break;
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    const bool clean = st.open_streams == 0 && st.open_readers == 0 && st.ring_slabs_free == st.ring_slabs_total &&
                       st.free_stream_slots == 256;
    if (!clean) printf("NOT CLEAN: open_streams=%llu open_readers=%llu slabs free/total=%llu/%llu free_stream_slots=%llu\n",
                       (unsigned long long)st.open_streams, (unsigned long long)st.open_readers, (unsigned long long)st.ring_slabs_free,
                       (unsigned long long)st.ring_slabs_total, (unsigned long long)st.free_stream_slots);
    if (g_inject) printf("injected device faults surfaced as DM_ECUDA: %ld\n", cuda_failed.load());
    printf("ops=%ld enomem=%ld followed=%ld launches=%llu committed=%llu mismatched=%llu ring_waits=%llu clean=%d failures=%d\n",
           ops.load(), enomem.load(), followed.load(), (unsigned long long)st.kernel_launches, (unsigned long long)st.blobs_committed,
           (unsigned long long)st.blobs_mismatched, (unsigned long long)st.ring_waits, (int)clean, failures.load());
    dm_engine_destroy(e);
    if (cas_dir && std::filesystem::exists(cas_dir)) {       // the disk tier may hold only files that hash to their own name
        long files = 0, bad = 0;
        for (auto &ent : std::filesystem::recursive_directory_iterator(cas_dir)) {
            if (!ent.is_regular_file()) continue;
            const std::string name = ent.path().filename().string();
            if (name.size() != 64) continue;                 // sidecars (.meta) and abandoned .part files are not blobs
            std::ifstream f(ent.path(), std::ios::binary);
            std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            uint8_t d[32];
            dmo_sha256(bytes.data(), bytes.size(), d);
            char hex[65];
            for (int i = 0; i < 32; ++i) snprintf(hex + 2 * i, 3, "%02x", d[i]);
            ++files;
            if (name != hex) { ++bad; fprintf(stderr, "disk tier file %s does not hash to its name\n", ent.path().c_str()); }
        }
        printf("disk tier: %ld blob files, %ld wrong\n", files, bad);
        if (bad) failures.fetch_add(1);
    }
    if (!failures.load() && !g_inject) driver_phase(verify_only);
    if (!failures.load() && !verify_only && !g_inject) destroy_with_open_handles(cas_dir);
    if (!failures.load() && !verify_only && !g_inject && !cas_dir) lost_bytes_are_sticky();
    if (!failures.load() && !verify_only && !g_inject && !cas_dir) bulk_replace_rewrites_cached_blobs_in_place();
    if (failures.load() || !clean) { printf("ENGINE SOAK FAILED\n"); return 1; }
    printf("ENGINE SOAK OK\n");
    return 0;
}
