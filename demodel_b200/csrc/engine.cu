// The blob hash-and-cache engine behind include/demodel_b200.h.
//
// Data path:
//
//   dm_stream_write ──memcpy──▶ pinned ring slab ──one H2D DMA per slab──▶ the blob's
//   (many threads)              (per stream)        (copy streams)         CAS extent in HBM
//                                                                               │
//   pump thread: gathers the streams with unhashed bytes into a job table       ▼
//   (one slab-sized job per stream, at most one in flight per stream) and  sha256_{deep,group,wide}
//   launches ONE multi-buffer SHA-256 kernel over them; up to 8 launches   (reads each byte once)
//   overlap on separate CUDA streams, each ordered after the DMAs by an
//   event.  Finished streams get their digest through mapped pinned memory,
//   are compared with the expected oid and published in the CAS index; spill
//   threads write published blobs to the disk tier with D2H copies on side
//   streams; readers (hits, followers of in-flight bodies) copy out through
//   pinned read-ahead windows.
//
// The bytes land at their final CAS address straight from the DMA, so the
// ring path costs HBM one write (DMA) + one read (hash) per blob byte.  The
// device-resident path (dm_ingest_device) fuses the copy into the hash kernel.
//
// Lock order (outer → inner): reader.mu → stream.mu → {arena_mu, e->mu, work_mu};
// slab_mu → work_mu; stripe / reader-stripe mutexes are leaves taken alone.
// The pump never holds work_mu while taking a stream mutex.
//
// Reference hooks served: cmd/demodel/start.go:201-204 (ingest) and
// start.go:197-200 (hit serving); see include/demodel_b200.h.
#include "../../include/demodel_b200.h"
#include "sha256_kernels.cuh"
#include "blobgen.h"
#include "host_util.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

namespace {

using dm::Arena;
using dm::add_interval;

thread_local std::string g_last_error;

int fail(int code, const char *what)
{
    g_last_error = what ? what : "";
    return code;
}
// cudaEventQuery's "not ready" also lands in the runtime's per-thread last-error slot, where the next
// `cudaGetLastError()` (the launch wrappers end in one) would find it and report a launch failure that never
// happened.  Every poll goes through here.
cudaError_t poll_event(cudaEvent_t ev)
{
    const cudaError_t q = cudaEventQuery(ev);
    if (q == cudaErrorNotReady) (void)cudaGetLastError();
    return q;
}

int fail_cuda(cudaError_t err, const char *where)
{
    g_last_error = std::string(where) + ": " + cudaGetErrorString(err);
    return DM_ECUDA;
}
#define CU_TRY(expr)                                                   \
    do {                                                               \
        cudaError_t cu_err_ = (expr);                                  \
        if (cu_err_ != cudaSuccess) return fail_cuda(cu_err_, #expr);  \
    } while (0)

constexpr uint64_t kAlign = 256;           // CAS extent granularity
constexpr uint64_t kMaxGrow = 256ull << 20;
constexpr int kCycles = 8;                 // concurrent hash launches, each on its own CUDA stream
constexpr int kSlabBatches = 64;           // groups of DMA'd slabs waiting for their copy events
constexpr int kStripes = 64;               // stream-table lock stripes
constexpr int kCopyStreams = 2;
constexpr size_t kBounceBytes = 4u << 20;
constexpr int kBounces = 48;               // 4 MiB each; readers borrow two as read-ahead windows
constexpr int kSpillThreads = 4;           // disk-tier writers (each double-buffers two bounce buffers)
constexpr int kBounceReserve = 2 * kSpillThreads + 2;          // never lent to windows: the spill thread and one-shot reads need some

inline uint64_t round_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct Digest {
    uint8_t b[32];
    bool operator==(const Digest &o) const { return memcmp(b, o.b, 32) == 0; }
};
struct DigestHash {
    size_t operator()(const Digest &d) const
    {
        uint64_t v;
        memcpy(&v, d.b, 8);   // SHA-256 output is already uniform
        return (size_t)v;
    }
};

void words_to_digest(const uint32_t *w, uint8_t out[32])
{
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(w[i] >> 24); out[4 * i + 1] = (uint8_t)(w[i] >> 16);
        out[4 * i + 2] = (uint8_t)(w[i] >> 8); out[4 * i + 3] = (uint8_t)w[i];
    }
}

std::string hex_of(const uint8_t *d, size_t n)
{
    static const char *hx = "0123456789abcdef";
    std::string s(2 * n, '0');
    for (size_t i = 0; i < n; ++i) { s[2 * i] = hx[d[i] >> 4]; s[2 * i + 1] = hx[d[i] & 15]; }
    return s;
}

struct Extent { uint64_t off, len; };      // byte range of the HBM arena

struct Blob {
    Digest digest;
    uint64_t size = 0;
    std::vector<Extent> extents;          // empty once evicted from HBM
    uint32_t readers = 0;
    uint64_t tick = 0;
    bool in_hbm = false;
    bool on_disk = false;
    bool spill_done = false;
    std::vector<std::pair<std::string, std::string>> meta;   // response headers to replay on a hit
};

struct Slab { uint8_t *host; uint8_t *dev; };   // dev: same slab of the device-side mirror (verify-only streams)

enum class St { Open, Finishing, Done, Aborted };

struct Stream {
    std::mutex mu;
    std::condition_variable cv;
    uint64_t id = 0;
    uint32_t slot = 0;
    bool has_expect = false;
    Digest expect{};
    std::vector<Extent> extents;
    uint64_t capacity = 0;     // sum of extents
    uint64_t received = 0;     // bytes accepted (count, any order)
    uint64_t dma_issued = 0;   // end of the contiguous run, starting at resume_base, whose H2D is enqueued
    uint64_t hash_issued = 0;  // bytes covered by launched jobs
    Slab *cur = nullptr;       // sequential cursor: stages [dma_issued, dma_issued + cur_fill)
    uint32_t cur_fill = 0;
    // Range parts (dm_stream_write_at): out-of-order pieces staged per part, DMA'd to their place in
    // the extent, and remembered as islands until the contiguous frontier reaches them.
    struct Part { uint64_t base; Slab *slab; uint32_t fill; };
    std::vector<Part> parts;
    std::map<uint64_t, uint64_t> islands;        // [start, end) DMA-enqueued beyond the frontier
    uint64_t resume_base = 0;                    // bytes hashed before this stream existed (checkpoint)
    std::map<uint64_t, uint64_t> prefix_cover;   // what of [0, resume_base) was re-supplied for caching
    bool ckpt_waiter = false;
    // Verify-only streams (DM_F_NO_HBM_CAS, or a blob that could never fit the arena): nothing is
    // retained.  Slabs are DMA'd to the device mirror of the ring and hashed from there, one slab per
    // job in arrival order; a slab returns to the ring when its job has run.
    bool verify_only = false;
    // A verify-only stream whose partly filled slab is recalled by ring back-pressure keeps the bytes past
    // the last whole block here (jobs hash whole blocks); they lead the stream's next slab.  carry_fill > 0
    // implies cur == nullptr.
    uint8_t carry[64];
    uint32_t carry_fill = 0;
    std::deque<std::pair<Slab *, uint32_t>> staged;
    std::vector<std::pair<std::string, std::string>> meta;   // dm_stream_set_meta
    uint32_t followers = 0;    // readers attached while the body is still arriving (request coalescing)
    uint32_t follow_reads = 0; // followers' copy-outs in flight: the extents must not be freed or handed over meanwhile
    bool completing = false;   // digest known, extents being handed to the index: followers wait for Done
    uint64_t size_hint = 0;
    bool window_out = false;   // acquire() window outstanding
    bool queued = false;       // in the pump's inbox / ready list (guarded by mu)
    bool final_issued = false;
    uint32_t jobs_inflight = 0;
    St st = St::Open;
    Digest digest{};
    int matched = 0;
    bool cuda_failed = false;  // a copy or launch for this stream failed: whatever digest comes back is not trusted
    int result = DM_OK;
    std::shared_ptr<Blob> blob;   // set at commit
};

struct Bounce;
struct Window { Bounce *b = nullptr; uint64_t off = 0, len = 0; bool pending = false; };

struct Reader {
    std::shared_ptr<Blob> blob;
    std::string disk_meta;        // sidecar text, disk-tier readers
    std::shared_ptr<Stream> follow;   // in-flight body this reader is coalesced onto (until it completes)
    int fd = -1;                  // disk tier
    uint64_t size = 0;
    std::mutex mu;                // a reader is normally one goroutine; this keeps misuse safe
    Window win[2];                // double-buffered read-ahead in pinned memory (HBM tier)
    bool tried_windows = false;
};

struct Cycle {
    bool busy = false;
    cudaEvent_t copy_ev[kCopyStreams]{};
    cudaEvent_t k_start{}, k_end{};
    cudaStream_t stream{};           // launches on different streams overlap on the GPU
    dm::HashJob *h_jobs = nullptr;   // pinned
    dm::HashJob *d_jobs = nullptr;
    uint32_t njobs = 0;
    bool deep = false;
    uint64_t bytes = 0;
    std::vector<std::shared_ptr<Stream>> streams;   // one entry per job
    std::vector<Slab *> job_slabs;                  // verify-only jobs: the ring slab to release at reap
    std::vector<uint8_t> is_final;
    cudaError_t err = cudaSuccess;                  // first failure while building or running this launch
};

struct SlabBatch {
    bool busy = false;
    cudaEvent_t ev[kCopyStreams]{};
    std::vector<Slab *> slabs;
};

struct Bounce { uint8_t *host = nullptr; cudaStream_t stream{}; };

}  // namespace

struct dm_engine {
    dm_config cfg{};
    std::string cas_dir;
    int device = 0;
    int sm_count = 148;
    int force_spw = 0;               // DM_FORCE_SPW: 1/2/4/8/16/32 streams per warp for every launch (tuning only)
    int variant_wide = dm::kDefaultWideVariant, variant_deep = dm::kDefaultDeepVariant;   // DM_KERNEL_VARIANT overrides (tuning only)

    cudaStream_t copy_stream[kCopyStreams]{};
    cudaStream_t ingest_stream{}, util_stream{};

    uint8_t *arena_base = nullptr;
    std::mutex arena_mu;
    Arena arena;

    uint8_t *ring = nullptr;
    uint8_t *dev_ring = nullptr;     // device mirror of the ring, allocated on first verify-only use
    std::vector<Slab> slab_store;
    std::mutex slab_mu;
    std::condition_variable slab_cv;
    std::vector<Slab *> slab_free;

    uint32_t *d_states = nullptr;
    uint32_t *h_digests = nullptr;   // mapped pinned, [max_streams][8]
    uint32_t *d_digests = nullptr;   // device alias of h_digests

    std::mutex mu;                   // blobs / readers / slots
    std::mutex stripe_mu[kStripes];  // stream table, striped by id: dm_stream_write never takes `mu`
    std::unordered_map<uint64_t, std::shared_ptr<Stream>> streams[kStripes];
    std::atomic<uint64_t> n_streams{0};
    std::atomic<bool> ring_starved{false};
    std::vector<uint32_t> free_slots;
    uint64_t next_id = 1;
    std::unordered_map<Digest, std::shared_ptr<Blob>, DigestHash> blobs;
    std::unordered_map<Digest, std::weak_ptr<Stream>, DigestHash> inflight;   // open streams by expected digest
    std::mutex reader_mu[kStripes];
    std::unordered_map<uint64_t, std::shared_ptr<Reader>> readers[kStripes];
    uint64_t tick = 0;

    std::mutex work_mu;              // pump inbox
    std::condition_variable work_cv;
    std::vector<std::shared_ptr<Stream>> dirty;
    std::vector<Slab *> pending_slabs;
    std::atomic<bool> stop{false};   // set under work_mu; the spill threads read it under spill_mu
    std::atomic<int> ring_waiters{0};   // writers blocked in slab_get(): the pump keeps recalling partial slabs meanwhile
    std::thread pump;
    Cycle cycles[kCycles];
    SlabBatch batches[kSlabBatches];
    uint32_t max_jobs = 0;

    std::mutex spill_mu;
    std::condition_variable spill_cv, spill_done_cv;
    std::deque<std::shared_ptr<Blob>> spill_q;
    std::vector<std::thread> spillers;   // kSpillThreads writers of the disk tier

    std::mutex bounce_mu;
    std::condition_variable bounce_cv;
    std::vector<Bounce *> bounce_free;
    std::vector<Bounce> bounce_store;

    std::mutex ckpt_mu;              // checkpoint state transfers: pinned staging + own stream, fully synchronous
    uint32_t *ckpt_pinned = nullptr;
    cudaStream_t ckpt_stream{};

    std::mutex ingest_mu;            // dm_ingest_device scratch
    uint32_t *ing_states = nullptr;
    uint32_t *ing_digests = nullptr;       // device
    dm::HashJob *ing_jobs_h = nullptr;     // pinned
    dm::HashJob *ing_jobs_d = nullptr;
    uint32_t *ing_digests_h = nullptr;     // pinned
    uint32_t ing_cap = 0;
    cudaEvent_t ing_ev0{}, ing_ev1{};

    // stats
    // the two counters bumped from caller threads sit on cache lines of their own (many writers / readers at once)
    alignas(64) std::atomic<uint64_t> st_ingested{0};
    alignas(64) std::atomic<uint64_t> st_served{0};
    alignas(64) std::atomic<uint64_t> st_hashed{0};
    std::atomic<uint64_t> st_committed{0}, st_mismatch{0};
    std::atomic<uint64_t> st_group{0};
    std::atomic<uint64_t> st_launches{0}, st_wide{0}, st_deep{0}, st_h2d{0}, st_d2h{0}, st_ring_waits{0};
    std::mutex stat_mu;
    double st_kernel_ms = 0.0;
};

namespace {

// ---- extents ---------------------------------------------------------------

// Visit the device segments covering [off, off+len) of a blob laid out over `ext`.
template <class F>
void for_segments(dm_engine *e, const std::vector<Extent> &ext, uint64_t off, uint64_t len, F &&fn)
{
    uint64_t base = 0;
    for (const Extent &x : ext) {
        if (len == 0) break;
        if (off < base + x.len) {
            const uint64_t in = off - base;
            const uint64_t n = std::min(len, x.len - in);
            fn(e->arena_base + x.off + in, n);
            off += n; len -= n;
        }
        base += x.len;
    }
}

// Device pointer of byte `off`, and how many bytes are contiguous from there.
uint8_t *seg_at(dm_engine *e, const std::vector<Extent> &ext, uint64_t off, uint64_t *contig)
{
    uint64_t base = 0;
    for (const Extent &x : ext) {
        if (off < base + x.len) { *contig = base + x.len - off; return e->arena_base + x.off + (off - base); }
        base += x.len;
    }
    *contig = 0;
    return nullptr;
}

void free_extents(dm_engine *e, std::vector<Extent> &ext)
{
    std::lock_guard<std::mutex> g(e->arena_mu);
    for (const Extent &x : ext) e->arena.release(x.off, x.len);
    ext.clear();
}

// Evict least-recently-used unreferenced blobs until `need` bytes could fit.
// Caller holds neither e->mu nor arena_mu.
bool evict_for(dm_engine *e, uint64_t need)
{
    for (;;) {
        std::shared_ptr<Blob> victim;
        std::vector<Extent> ext;
        {
            std::lock_guard<std::mutex> g(e->mu);
            for (auto &kv : e->blobs) {
                Blob *b = kv.second.get();
                if (!b->in_hbm || b->readers) continue;
                if (!e->cas_dir.empty() && !b->spill_done) continue;   // not yet safe on disk
                if (!victim || b->tick < victim->tick) victim = kv.second;
            }
            if (!victim) return false;
            victim->in_hbm = false;
            ext.swap(victim->extents);          // taken under the lock: a re-publish may install new ones at once
            if (!victim->on_disk) e->blobs.erase(victim->digest);
        }
        free_extents(e, ext);
        std::lock_guard<std::mutex> g(e->arena_mu);
        uint64_t off;
        if (e->arena.alloc(need, &off)) { e->arena.release(off, need); return true; }
    }
}

bool arena_alloc(dm_engine *e, uint64_t len, Extent *out)
{
    len = round_up(std::max<uint64_t>(len, 1), kAlign);
    for (int attempt = 0; attempt < 4; ++attempt) {      // another thread may take what an eviction freed
        {
            std::lock_guard<std::mutex> g(e->arena_mu);
            uint64_t off;
            if (e->arena.alloc(len, &off)) { out->off = off; out->len = len; return true; }
        }
        if (!evict_for(e, len)) return false;
    }
    return false;
}

// Make sure the stream's extents cover `need` bytes.  Stream mutex held.
int ensure_capacity(dm_engine *e, Stream *s, uint64_t need)
{
    while (s->capacity < need) {
        uint64_t want = need - s->capacity;
        if (!s->extents.empty()) {   // unknown / exceeded size: grow geometrically
            const uint64_t grow = std::min<uint64_t>(std::max<uint64_t>(s->capacity, e->cfg.slab_bytes), kMaxGrow);
            want = std::max(want, grow);
        }
        Extent x;
        if (!arena_alloc(e, want, &x)) {
            if (!arena_alloc(e, need - s->capacity, &x)) return fail(DM_ENOMEM, "HBM CAS arena exhausted");
        }
        s->extents.push_back(x);
        s->capacity += x.len;
    }
    return DM_OK;
}

// ---- ring slabs --------------------------------------------------------------

Slab *slab_get(dm_engine *e)
{
    std::unique_lock<std::mutex> g(e->slab_mu);
    if (e->slab_free.empty() && !e->stop) {
        e->st_ring_waits++;
        // Every slab is out: some are only partly filled and held by streams waiting for their next
        // bytes.  Ask the pump to DMA those early so they recycle (otherwise more live streams than
        // slabs would starve — or, with one thread driving many streams, deadlock).
        { std::lock_guard<std::mutex> gw(e->work_mu); e->ring_starved.store(true); }   // under the pump's mutex: no lost wake-up
        e->work_cv.notify_one();
    }
    e->ring_waiters++;
    e->slab_cv.wait(g, [&] { return !e->slab_free.empty() || e->stop; });
    e->ring_waiters--;
    if (e->slab_free.empty()) return nullptr;
    Slab *s = e->slab_free.back();
    e->slab_free.pop_back();
    return s;
}

void slab_put(dm_engine *e, Slab *s)
{
    {
        std::lock_guard<std::mutex> g(e->slab_mu);
        e->slab_free.push_back(s);
    }
    e->slab_cv.notify_one();
}

// Give the stream a fresh slab.  The stream mutex is dropped while waiting for
// ring back-pressure: the pump needs it to build jobs, and only the pump's
// reaping frees slabs.
int take_slab(dm_engine *e, Stream *s, std::unique_lock<std::mutex> &g)
{
    g.unlock();
    Slab *fresh = slab_get(e);
    g.lock();
    if (!fresh) return fail(DM_ESTATE, "engine stopping");
    if (s->st != St::Open) { slab_put(e, fresh); return fail(DM_ESTATE, "stream closed while waiting for the ring"); }
    if (s->cur) { slab_put(e, fresh); return DM_OK; }
    s->cur = fresh;
    s->cur_fill = s->carry_fill;
    if (s->carry_fill) { memcpy(fresh->host, s->carry, s->carry_fill); s->carry_fill = 0; }
    return DM_OK;
}

// Tell the pump this stream has new DMA'd bytes (or is finishing).  Stream mutex held.
void mark_dirty(dm_engine *e, const std::shared_ptr<Stream> &sp, Slab *submitted)
{
    const bool enqueue = !sp->queued;
    sp->queued = true;
    {
        std::lock_guard<std::mutex> g(e->work_mu);
        if (submitted) e->pending_slabs.push_back(submitted);
        if (enqueue) e->dirty.push_back(sp);
    }
    e->work_cv.notify_one();
}

// DMA `n` staged bytes to blob offset `base`.  Stream mutex held.  The slab is handed to the pump
// (released once its copy event completes) on success, returned to the ring on failure.
int dma_range(dm_engine *e, const std::shared_ptr<Stream> &sp, Slab *slab, uint64_t base, uint32_t n)
{
    Stream *s = sp.get();
    cudaSetDevice(e->device);       // the caller may be any OS thread (cgo)
    int rc = ensure_capacity(e, s, base + n);
    if (rc != DM_OK) { slab_put(e, slab); return rc; }
    cudaStream_t cs = e->copy_stream[s->id % kCopyStreams];
    const uint8_t *src = slab->host;
    cudaError_t err = cudaSuccess;
    for_segments(e, s->extents, base, n, [&](uint8_t *dev, uint64_t len) {
        if (err == cudaSuccess) err = cudaMemcpyAsync(dev, src, len, cudaMemcpyHostToDevice, cs);
        src += len;
    });
    if (err != cudaSuccess) { s->cuda_failed = true; slab_put(e, slab); return fail_cuda(err, "cudaMemcpyAsync(H2D slab)"); }
    e->st_h2d += n;
    return DM_OK;
}

// The contiguous frontier swallows islands that now touch it.  Stream mutex held.
void absorb_islands(Stream *s)
{
    if (s->cur_fill) return;            // staged sequential bytes sit between the frontier and any island
    auto it = s->islands.begin();
    while (it != s->islands.end() && it->first <= s->dma_issued) {
        s->dma_issued = std::max(s->dma_issued, it->second);
        it = s->islands.erase(it);
    }
}

// DMA the stream's sequential slab to its place in the blob.  Stream mutex held.
int submit_slab(dm_engine *e, const std::shared_ptr<Stream> &sp)
{
    Stream *s = sp.get();
    if (!s->cur) return DM_OK;
    Slab *slab = s->cur;
    const uint32_t n = s->cur_fill;
    s->cur = nullptr; s->cur_fill = 0;
    if (n == 0) { slab_put(e, slab); return DM_OK; }
    e->st_ingested.fetch_add(n, std::memory_order_relaxed);       // per slab, not per write: one shared line, many writer threads
    if (s->verify_only) {
        cudaSetDevice(e->device);
        cudaError_t err = cudaMemcpyAsync(slab->dev, slab->host, n, cudaMemcpyHostToDevice, e->copy_stream[s->id % kCopyStreams]);
        if (err != cudaSuccess) { slab_put(e, slab); return fail_cuda(err, "cudaMemcpyAsync(H2D slab)"); }
        e->st_h2d += n;
        s->staged.emplace_back(slab, n);
        s->dma_issued += n;
        mark_dirty(e, sp, nullptr);       // the slab stays out of the ring until its job has run
        return DM_OK;
    }
    int rc = dma_range(e, sp, slab, s->dma_issued, n);
    if (rc != DM_OK) return rc;
    s->dma_issued += n;
    absorb_islands(s);
    mark_dirty(e, sp, slab);
    if (s->followers) s->cv.notify_all();
    return DM_OK;
}

// DMA one range part.  Stream mutex held; invalidates indices into s->parts.
int submit_part(dm_engine *e, const std::shared_ptr<Stream> &sp, size_t idx)
{
    Stream *s = sp.get();
    Stream::Part pt = s->parts[idx];
    s->parts.erase(s->parts.begin() + (long)idx);
    if (pt.fill == 0) { slab_put(e, pt.slab); return DM_OK; }
    e->st_ingested.fetch_add(pt.fill, std::memory_order_relaxed);
    int rc = dma_range(e, sp, pt.slab, pt.base, pt.fill);
    if (rc != DM_OK) return rc;
    if (pt.base + pt.fill <= s->resume_base) add_interval(s->prefix_cover, pt.base, pt.base + pt.fill);
    else add_interval(s->islands, pt.base, pt.base + pt.fill);
    absorb_islands(s);
    mark_dirty(e, sp, pt.slab);
    if (s->followers) s->cv.notify_all();
    return DM_OK;
}

// Would [off, off+len) collide with bytes this stream already holds?  Stream mutex held.
bool range_taken(const Stream *s, uint64_t off, uint64_t len, const Stream::Part *self)
{
    const uint64_t end = off + len;
    auto hits = [&](uint64_t lo, uint64_t hi) { return lo < end && off < hi; };
    if (hits(s->resume_base, s->dma_issued + s->cur_fill) && !(self == nullptr && off == s->dma_issued + s->cur_fill)) return true;
    auto in_map = [&](const std::map<uint64_t, uint64_t> &m) {
        auto it = m.upper_bound(off);
        if (it != m.begin() && std::prev(it)->second > off) return true;
        return it != m.end() && it->first < end;
    };
    if (in_map(s->islands) || in_map(s->prefix_cover)) return true;
    for (const Stream::Part &p : s->parts)
        if (&p != self && hits(p.base, p.base + p.fill)) return true;
    return false;
}

// ---- CAS commit --------------------------------------------------------------

std::string json_quote(const std::string &v)
{
    std::string o = "\"";
    for (unsigned char c : v) {
        if (c == '"' || c == '\\') { o += '\\'; o += (char)c; }
        else if (c < 0x20) { char t[8]; snprintf(t, sizeof t, "\\u%04x", c); o += t; }
        else o += (char)c;
    }
    return o + "\"";
}

// The sidecar: what a hit needs besides the bytes.  The digest is over the
// identity-encoded body, which is what HF LFS oids and OCI layer digests are defined on.
std::string sidecar_json(const Blob &b)
{
    std::string o = "{\"digest\":\"sha256:" + hex_of(b.digest.b, 32) + "\",\"size\":" + std::to_string(b.size) +
                    ",\"encoding\":\"identity\",\"engine\":\"demodel_b200\",\"abi\":" + std::to_string(DM_ABI_VERSION) +
                    ",\"headers\":{";
    for (size_t i = 0; i < b.meta.size(); ++i)
        o += (i ? "," : "") + json_quote(b.meta[i].first) + ":" + json_quote(b.meta[i].second);
    return o + "}}\n";
}

void write_sidecar(const std::string &path, const Blob &b)
{
    FILE *f = fopen(path.c_str(), "w");
    if (!f) return;
    const std::string j = sidecar_json(b);
    fwrite(j.data(), 1, j.size(), f);
    fclose(f);
}

std::string blob_path(const dm_engine *e, const uint8_t d[32])
{
    const std::string hx = hex_of(d, 32);
    return e->cas_dir + "/blobs/sha256/" + hx.substr(0, 2) + "/" + hx;
}

// Publish a verified blob.  Returns the blob that now owns the digest (an
// earlier copy wins; the new extents are then released).
std::shared_ptr<Blob> publish(dm_engine *e, const Digest &d, uint64_t size, std::vector<Extent> &ext,
                              std::vector<std::pair<std::string, std::string>> *meta = nullptr)
{
    // trim the last extent to the bytes actually held
    uint64_t keep = round_up(std::max<uint64_t>(size, 1), kAlign), base = 0;
    std::vector<Extent> kept;
    {
        std::lock_guard<std::mutex> g(e->arena_mu);
        for (Extent &x : ext) {
            if (base >= keep) { e->arena.release(x.off, x.len); }
            else if (base + x.len > keep) {
                const uint64_t k = keep - base;
                e->arena.release(x.off + k, x.len - k);
                kept.push_back({x.off, k});
            } else kept.push_back(x);
            base += x.len;
        }
    }
    ext.clear();
    std::shared_ptr<Blob> b;
    bool fresh = false;
    {
        std::lock_guard<std::mutex> g(e->mu);
        auto it = e->blobs.find(d);
        if (it != e->blobs.end() && it->second->in_hbm) {
            b = it->second;
            b->tick = ++e->tick;
        } else if (it != e->blobs.end()) {   // known on disk only: re-home into HBM
            b = it->second;
            b->extents = kept; kept.clear();
            b->in_hbm = true; b->tick = ++e->tick;
        } else {
            b = std::make_shared<Blob>();
            b->digest = d; b->size = size; b->extents = kept; kept.clear();
            b->in_hbm = true; b->tick = ++e->tick;
            if (meta) b->meta.swap(*meta);
            e->blobs[d] = b;
            fresh = true;
        }
    }
    if (!kept.empty()) free_extents(e, kept);
    e->st_committed++;
    if (fresh && !e->cas_dir.empty()) {
        {
            std::lock_guard<std::mutex> g(e->spill_mu);
            e->spill_q.push_back(b);
        }
        e->spill_cv.notify_all();
    }
    return b;
}

// Batch form of publish() for dm_ingest_device: every blob has exactly one right-sized extent, so
// the index is updated under ONE lock and the arena under one more (150 k blobs per call otherwise
// spend longer in lock traffic than in the kernel).
struct Verified { Digest d; uint64_t size; Extent x; };
void publish_many(dm_engine *e, const std::vector<Verified> &items)
{
    std::vector<Extent> to_free;
    std::vector<std::shared_ptr<Blob>> to_spill;
    {
        std::lock_guard<std::mutex> g(e->mu);
        for (const Verified &v : items) {
            auto it = e->blobs.find(v.d);
            if (it != e->blobs.end() && it->second->in_hbm) {          // an earlier copy wins
                it->second->tick = ++e->tick;
                to_free.push_back(v.x);
            } else if (it != e->blobs.end()) {                         // on disk only: re-home
                Blob *b = it->second.get();
                b->extents.assign(1, v.x);
                b->in_hbm = true; b->tick = ++e->tick;
            } else {
                auto b = std::make_shared<Blob>();
                b->digest = v.d; b->size = v.size; b->extents.assign(1, v.x);
                b->in_hbm = true; b->tick = ++e->tick;
                e->blobs.emplace(v.d, b);
                if (!e->cas_dir.empty()) to_spill.push_back(b);
            }
        }
    }
    e->st_committed += items.size();
    if (!to_free.empty()) {
        std::lock_guard<std::mutex> g(e->arena_mu);
        for (const Extent &x : to_free) e->arena.release(x.off, x.len);
    }
    if (!to_spill.empty()) {
        {
            std::lock_guard<std::mutex> g(e->spill_mu);
            for (auto &b : to_spill) e->spill_q.push_back(b);
        }
        e->spill_cv.notify_all();
    }
}

// Batch eviction (DM_ING_REPLACE): drop the HBM copies of these digests under one lock each way.
void evict_many(dm_engine *e, const uint8_t *digests, uint32_t n)
{
    std::vector<Extent> ext;
    {
        std::lock_guard<std::mutex> g(e->mu);
        for (uint32_t i = 0; i < n; ++i) {
            Digest d;
            memcpy(d.b, digests + 32ull * i, 32);
            auto it = e->blobs.find(d);
            if (it == e->blobs.end() || !it->second->in_hbm || it->second->readers) continue;
            it->second->in_hbm = false;
            ext.insert(ext.end(), it->second->extents.begin(), it->second->extents.end());
            it->second->extents.clear();            // under the lock (see evict_for)
            if (!it->second->on_disk) e->blobs.erase(it);
        }
    }
    free_extents(e, ext);
}

// ---- pump ----------------------------------------------------------------------

// Before a stream's extents are freed or change owner: let followers' in-flight copy-outs finish.
void wait_follow_reads(Stream *s, std::unique_lock<std::mutex> &g)
{
    s->cv.wait(g, [&] { return s->follow_reads == 0; });
}

void complete_stream(dm_engine *e, const std::shared_ptr<Stream> &sp, const uint32_t *words)
{
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    if (s->st == St::Aborted) return;
    words_to_digest(words, s->digest.b);
    s->matched = (!s->has_expect || s->digest == s->expect) ? 1 : 0;
    if (s->cuda_failed) { s->matched = 0; memset(s->digest.b, 0, 32); }     // never publish under a digest the device may not have produced
    s->completing = true;                       // no new follower copy-out starts past this point
    wait_follow_reads(s, g);
    std::vector<Extent> ext;
    ext.swap(s->extents);
    s->capacity = 0;
    const uint64_t size = s->dma_issued;
    const Digest d = s->digest;
    const int matched = s->matched;
    // a resumed stream is cacheable only if the already-hashed prefix was re-supplied
    const bool whole = s->resume_base == 0 ||
                       (s->prefix_cover.size() == 1 && s->prefix_cover.begin()->first == 0 &&
                        s->prefix_cover.begin()->second >= s->resume_base);
    g.unlock();
    std::shared_ptr<Blob> b;
    std::vector<std::pair<std::string, std::string>> meta;
    g.lock(); meta.swap(s->meta); g.unlock();
    if (matched && whole && !s->verify_only) b = publish(e, d, size, ext, &meta);
    else { free_extents(e, ext); if (!matched) e->st_mismatch++; }
    g.lock();
    s->blob = b;
    s->st = St::Done;
    g.unlock();
    s->cv.notify_all();
}

void reap_cycle(dm_engine *e, Cycle &c)
{
    float ms = 0.f;
    if (c.njobs) {
        cudaEventElapsedTime(&ms, c.k_start, c.k_end);
        std::lock_guard<std::mutex> g(e->stat_mu);
        e->st_kernel_ms += ms;
    }
    e->st_hashed += c.bytes;
    for (Slab *sl : c.job_slabs) if (sl) slab_put(e, sl);
    c.job_slabs.clear();
    for (size_t i = 0; i < c.streams.size(); ++i) {
        std::shared_ptr<Stream> &sp = c.streams[i];
        bool free_now = false, wake = false;
        {
            std::lock_guard<std::mutex> g(sp->mu);
            if (c.err != cudaSuccess) sp->cuda_failed = true;
            sp->jobs_inflight--;
            free_now = sp->st == St::Aborted && sp->jobs_inflight == 0;
            wake = sp->ckpt_waiter;
        }
        if (wake) sp->cv.notify_all();
        if (free_now) {
            // slabs written after this job was built may still be landing in the extent (see dm_stream_abort)
            cudaStreamSynchronize(e->copy_stream[sp->id % kCopyStreams]);
            { std::unique_lock<std::mutex> g(sp->mu); wait_follow_reads(sp.get(), g); }
            free_extents(e, sp->extents);
            std::lock_guard<std::mutex> g(e->mu);
            e->free_slots.push_back(sp->slot);
        } else if (c.is_final[i]) {
            complete_stream(e, sp, e->h_digests + 8ull * sp->slot);
        }
    }
    c.streams.clear(); c.is_final.clear();
    c.njobs = 0; c.bytes = 0; c.busy = false; c.err = cudaSuccess;
}

// Build one job per eligible ready stream and launch ONE multi-buffer kernel
// over them on this cycle's CUDA stream.  Streams that still have unhashed
// bytes afterwards (or a job in flight) stay in `ready`.
bool run_cycle(dm_engine *e, Cycle &c, std::vector<std::shared_ptr<Stream>> &ready)
{
    c.njobs = 0; c.bytes = 0;
    // One slab per job: a launch lasts as long as its longest lane, so lanes are kept the same
    // length (streams holding more simply go again in the next launch, which overlaps this one).
    const uint64_t quantum = (uint64_t)e->cfg.slab_bytes;
    std::vector<std::shared_ptr<Stream>> again;
    for (auto &sp : ready) {
        Stream *s = sp.get();
        std::lock_guard<std::mutex> g(s->mu);
        if (s->st == St::Aborted || s->st == St::Done || s->final_issued) { s->queued = false; continue; }
        // one job per stream in flight: its next job chains on the state this one writes
        if (s->jobs_inflight || c.njobs >= e->max_jobs) { again.push_back(sp); continue; }
        const bool finishing = s->st == St::Finishing;
        uint64_t n = finishing ? (s->dma_issued - s->hash_issued) : ((s->dma_issued - s->hash_issued) & ~63ull);
        if (!finishing && n == 0) { s->queued = false; continue; }
        uint64_t contig = 0;
        uint8_t *src = nullptr;
        bool final = finishing;
        Slab *job_slab = nullptr;
        if (s->verify_only) {
            if (!s->staged.empty()) {
                job_slab = s->staged.front().first;
                n = s->staged.front().second;               // whole slab; only the last may hold a partial block
                src = job_slab->dev;
                s->staged.pop_front();
                final = finishing && s->staged.empty();
                if (!final && (n & 63)) { final = false; }   // cannot happen: mid-stream slabs are full
            } else if (!finishing) { s->queued = false; continue; }
            else n = 0;
        } else {
            src = n ? seg_at(e, s->extents, s->hash_issued, &contig) : nullptr;
            if (n > contig && n) { n = contig; final = false; }          // stop at the extent boundary
            if (n > quantum) { n = quantum; final = false; }
        }
        dm::HashJob &jb = c.h_jobs[c.njobs++];
        jb.src = src; jb.dst = nullptr; jb.nbytes = n; jb.total_len = s->dma_issued; jb.slot = s->slot;
        jb.flags = (s->hash_issued == 0 ? dm::JOB_INIT : 0u) | (final ? dm::JOB_FINAL : 0u);
        jb.one = 1; jb.pad_ = 0;
        s->hash_issued += n;
        s->jobs_inflight++;
        if (final) s->final_issued = true;
        c.bytes += n;
        c.streams.push_back(sp);
        c.is_final.push_back(final ? 1 : 0);
        c.job_slabs.push_back(job_slab);
        if (!final && (finishing || !s->staged.empty() || (!s->verify_only && ((s->dma_issued - s->hash_issued) & ~63ull)))) again.push_back(sp);
        else s->queued = false;
    }
    ready.swap(again);
    if (c.njobs == 0) return false;

    // Everything whose DMA was enqueued before this point is covered by these events.
    for (int i = 0; i < kCopyStreams; ++i) {
        cudaEventRecord(c.copy_ev[i], e->copy_stream[i]);
        cudaStreamWaitEvent(c.stream, c.copy_ev[i], 0);
    }
    // Launches overlap on the GPU, so what decides the kernel shape is how many jobs will be
    // co-resident (this launch + those still running), not the size of this launch alone.
    uint32_t resident = c.njobs;
    for (const Cycle &o : e->cycles) if (o.busy) resident += o.njobs;
    const int spw = e->force_spw ? e->force_spw : dm::streams_per_warp_for(resident);
    c.deep = spw == 1;
    if (spw > 1) {
        // lanes of a warp run in lock step: keep neighbours the same length
        std::vector<uint32_t> order(c.njobs);
        for (uint32_t i = 0; i < c.njobs; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(),
                         [&](uint32_t a, uint32_t b) { return c.h_jobs[a].nbytes > c.h_jobs[b].nbytes; });
        std::vector<dm::HashJob> tmp(c.h_jobs, c.h_jobs + c.njobs);
        std::vector<std::shared_ptr<Stream>> st2(c.njobs);
        std::vector<uint8_t> fin2(c.njobs);
        std::vector<Slab *> sl2(c.njobs);
        for (uint32_t i = 0; i < c.njobs; ++i) { c.h_jobs[i] = tmp[order[i]]; st2[i] = c.streams[order[i]]; fin2[i] = c.is_final[order[i]]; sl2[i] = c.job_slabs[order[i]]; }
        c.streams.swap(st2); c.is_final.swap(fin2); c.job_slabs.swap(sl2);
    }
    // A failure anywhere here (or reported later by the end event) marks every stream of the launch:
    // their verdict becomes "not matched" and nothing is published (reap_cycle / complete_stream).
    (void)cudaGetLastError();           // nothing stale may be mistaken for this launch's result
    auto note = [&](cudaError_t r) { if (r != cudaSuccess && r != cudaErrorNotReady && c.err == cudaSuccess) c.err = r; };
    note(cudaMemcpyAsync(c.d_jobs, c.h_jobs, sizeof(dm::HashJob) * c.njobs, cudaMemcpyHostToDevice, c.stream));
    note(cudaEventRecord(c.k_start, c.stream));
    if (c.err != cudaSuccess) { /* the job table may not be on the device: launching would run stale jobs */ }
    else if (spw == 1) { note(dm::launch_sha256_deep(c.d_jobs, c.njobs, e->d_states, e->d_digests, c.stream, e->variant_deep)); e->st_deep++; }
    else if (spw == 32) { note(dm::launch_sha256_wide(c.d_jobs, c.njobs, e->d_states, e->d_digests, c.stream, e->variant_wide)); e->st_wide++; }
    else { note(dm::launch_sha256_group(c.d_jobs, c.njobs, e->d_states, e->d_digests, c.stream, spw, e->variant_deep)); e->st_group++; }
    e->st_launches++;
    note(cudaEventRecord(c.k_end, c.stream));
    c.busy = true;
    return true;
}

// Ring back-pressure relief (pump thread): DMA the partly filled slabs of open streams - the sequential
// one and those of range parts (a client fetching one blob as P parallel ranges holds P of them) - so
// they return to the ring.  try_lock only: a stream busy in a write keeps its slabs this round.
void flush_partial_slabs(dm_engine *e)
{
    std::vector<std::shared_ptr<Stream>> all;
    for (int k = 0; k < kStripes; ++k) {
        std::lock_guard<std::mutex> g(e->stripe_mu[k]);
        for (auto &kv : e->streams[k]) all.push_back(kv.second);
    }
    for (auto &sp : all) {
        Stream *s = sp.get();
        std::unique_lock<std::mutex> g(s->mu, std::try_to_lock);
        if (!g.owns_lock() || s->st != St::Open || s->window_out) continue;
        // a part that was sent early simply becomes an island; the range continues in a fresh part
        for (size_t i = s->parts.size(); i-- > 0;)
            if (s->parts[i].fill) submit_part(e, sp, i);
        if (!s->cur || s->cur_fill == 0) continue;
        if (s->verify_only && (s->cur_fill & 63)) {              // slab-by-slab hashing needs whole blocks:
            const uint32_t whole = s->cur_fill & ~63u;           // send those, keep the tail in the stream
            s->carry_fill = s->cur_fill - whole;
            memcpy(s->carry, s->cur->host + whole, s->carry_fill);
            s->cur_fill = whole;                                 // (0 whole blocks: submit_slab just returns the slab)
        }
        submit_slab(e, sp);
    }
}

// The pump owns all launch decisions.  Policy: at most one JOB PER STREAM in
// flight (its next job chains on the state the running one writes), but up to
// kCycles LAUNCHES in flight, each on its own CUDA stream.  A launch costs its
// longest lane however few lanes it has, so a launch that caught only a few
// early streams must not hold the others back: they go out in the next launch
// and overlap with it on the GPU (a deep launch occupies one sub-partition per
// job).  When every launch slot is busy the ready set simply accumulates.
void pump_main(dm_engine *e)
{
    cudaSetDevice(e->device);
    int n_inflight = 0;
    int b_head = 0, b_tail = 0, b_live = 0;
    std::vector<std::shared_ptr<Stream>> ready, inbox;
    std::vector<Slab *> slabs;
    int starve_ticks = 0;
    bool retry_ready = false;       // ready streams blocked only by their own in-flight job
    bool launched = false;
    for (;;) {
        // 1. ring slabs whose DMA has completed go back to the writers
        while (b_live) {
            SlabBatch &b = e->batches[b_tail];
            bool done = true;
            for (int i = 0; i < kCopyStreams; ++i) done = done && poll_event(b.ev[i]) == cudaSuccess;
            if (!done) break;
            for (Slab *sl : b.slabs) slab_put(e, sl);
            b.slabs.clear(); b.busy = false;
            b_tail = (b_tail + 1) % kSlabBatches; --b_live;
        }
        // a recall can miss slabs (stream busy in a write, window lent out): repeat while writers wait
        if (e->ring_starved.exchange(false) || (e->ring_waiters.load() > 0 && ++starve_ticks >= 16)) {
            starve_ticks = 0;
            flush_partial_slabs(e);
        }
        // 2. finished hash launches (any order)
        bool reaped = false;
        for (Cycle &c : e->cycles)
            if (c.busy) {
                const cudaError_t q = poll_event(c.k_end);
                if (q == cudaErrorNotReady) continue;
                if (q != cudaSuccess && c.err == cudaSuccess) c.err = q;       // a faulted launch must end, not hang its streams
                reap_cycle(e, c); --n_inflight; reaped = true;
            }
        // 3. inbox.  Sleep unless the previous pass launched something (more may be launchable).
        bool stopping;
        {
            std::unique_lock<std::mutex> g(e->work_mu);
            const bool idle = !n_inflight && !b_live && ready.empty() && slabs.empty();
            if (e->dirty.empty() && e->pending_slabs.empty() && !e->stop && !launched && !reaped) {
                if (idle && e->ring_waiters.load() == 0)
                    e->work_cv.wait(g, [&] { return !e->dirty.empty() || !e->pending_slabs.empty() || e->stop || e->ring_starved.load(); });
                else e->work_cv.wait_for(g, std::chrono::microseconds(40));
            }
            stopping = e->stop;
            inbox.swap(e->dirty);
            if (slabs.empty()) slabs.swap(e->pending_slabs);
            else { slabs.insert(slabs.end(), e->pending_slabs.begin(), e->pending_slabs.end()); e->pending_slabs.clear(); }
        }
        launched = false;
        const bool fresh = !inbox.empty();
        for (auto &sp : inbox) ready.push_back(sp);
        inbox.clear();
        // 4. tag the newly DMA'd slabs with copy events
        if (!slabs.empty() && b_live < kSlabBatches) {
            SlabBatch &b = e->batches[b_head];
            b.slabs.swap(slabs);
            for (int i = 0; i < kCopyStreams; ++i) cudaEventRecord(b.ev[i], e->copy_stream[i]);
            b.busy = true;
            b_head = (b_head + 1) % kSlabBatches; ++b_live;
        }
        // 5. launch on a free slot.  With launches already running, let the ready set
        //    build up to a worthwhile size first (they will all fit in one launch anyway).
        if (!ready.empty() && n_inflight < kCycles && (fresh || reaped || !retry_ready)) {
            const uint64_t open_now = e->n_streams;
            const bool worthwhile = n_inflight == 0 || ready.size() * 8 >= open_now || ready.size() >= 4096;
            if (worthwhile) {
                for (Cycle &c : e->cycles) {
                    if (c.busy) continue;
                    const size_t before = ready.size();
                    if (run_cycle(e, c, ready)) { ++n_inflight; retry_ready = false; launched = true; }
                    else retry_ready = !ready.empty() && ready.size() == before;   // all blocked on their own jobs
                    break;
                }
            }
        }
        if (stopping && !n_inflight && !b_live && ready.empty() && slabs.empty()) {
            std::lock_guard<std::mutex> g(e->work_mu);
            if (e->dirty.empty() && e->pending_slabs.empty()) break;
        }
    }
}

// ---- disk tier -----------------------------------------------------------------

Bounce *bounce_get(dm_engine *e)
{
    std::unique_lock<std::mutex> g(e->bounce_mu);
    e->bounce_cv.wait(g, [&] { return !e->bounce_free.empty(); });
    Bounce *b = e->bounce_free.back();
    e->bounce_free.pop_back();
    return b;
}
Bounce *bounce_try_get(dm_engine *e)      // for long-lived borrowers: leaves a reserve
{
    std::lock_guard<std::mutex> g(e->bounce_mu);
    if ((int)e->bounce_free.size() <= kBounceReserve) return nullptr;
    Bounce *b = e->bounce_free.back();
    e->bounce_free.pop_back();
    return b;
}
void bounce_put(dm_engine *e, Bounce *b)
{
    { std::lock_guard<std::mutex> g(e->bounce_mu); e->bounce_free.push_back(b); }
    e->bounce_cv.notify_one();
}

void mkdirs(const std::string &path)
{
    for (size_t i = 1; i < path.size(); ++i)
        if (path[i] == '/') { std::string p = path.substr(0, i); mkdir(p.c_str(), 0755); }
}

bool spill_one(dm_engine *e, Blob *b)
{
    const std::string path = blob_path(e, b->digest.b), tmp = path + ".part";
    mkdirs(path);
    int fd = open(tmp.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
    if (fd < 0) return false;
    // two pinned buffers: the D2H of piece k+1 runs while piece k is written to the file
    Bounce *bn[2] = {bounce_get(e), bounce_get(e)};
    uint64_t piece_len[2] = {0, 0};
    bool ok = true;
    auto start_piece = [&](int slot, uint64_t off) {
        const uint64_t n = std::min<uint64_t>(kBounceBytes, b->size - off);
        uint8_t *dst = bn[slot]->host;
        cudaError_t err = cudaSuccess;
        for_segments(e, b->extents, off, n, [&](uint8_t *dev, uint64_t len) {
            if (err == cudaSuccess) err = cudaMemcpyAsync(dst, dev, len, cudaMemcpyDeviceToHost, bn[slot]->stream);
            dst += len;
        });
        piece_len[slot] = n;
        return err == cudaSuccess;
    };
    uint64_t issued = 0, written = 0;
    int cur = 0;
    if (b->size) { ok = start_piece(0, 0); issued = piece_len[0]; }
    while (ok && written < b->size) {
        if (issued < b->size) { ok = start_piece(cur ^ 1, issued); issued += piece_len[cur ^ 1]; }
        if (cudaStreamSynchronize(bn[cur]->stream) != cudaSuccess) { ok = false; break; }
        const uint64_t n = piece_len[cur];
        e->st_d2h += n;
        uint64_t w = 0;
        while (w < n) {
            ssize_t r = write(fd, bn[cur]->host + w, n - w);
            if (r < 0) { if (errno == EINTR) continue; ok = false; break; }
            w += (uint64_t)r;
        }
        written += n;
        cur ^= 1;
    }
    cudaStreamSynchronize(bn[0]->stream);
    cudaStreamSynchronize(bn[1]->stream);
    bounce_put(e, bn[0]);
    bounce_put(e, bn[1]);
    close(fd);
    if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;
    if (ok) write_sidecar(path + ".meta", *b);
    else unlink(tmp.c_str());
    return ok;
}

void spill_main(dm_engine *e)
{
    cudaSetDevice(e->device);
    for (;;) {
        std::shared_ptr<Blob> b;
        {
            std::unique_lock<std::mutex> g(e->spill_mu);
            e->spill_cv.wait(g, [&] { return !e->spill_q.empty() || e->stop; });
            if (e->spill_q.empty()) break;
            b = e->spill_q.front();
            e->spill_q.pop_front();
        }
        bool have;
        {
            std::lock_guard<std::mutex> g(e->mu);
            have = b->in_hbm;
            if (have) b->readers++;        // pin against eviction while copying out
        }
        bool ok = have && spill_one(e, b.get());
        {
            std::lock_guard<std::mutex> g(e->mu);
            if (have) b->readers--;
            b->on_disk = ok;
            b->spill_done = true;
        }
        {
            std::lock_guard<std::mutex> g(e->spill_mu);
        }
        e->spill_done_cv.notify_all();
    }
}

// The device mirror of the ring exists from the start with DM_F_NO_HBM_CAS, otherwise it is
// allocated the first time a blob too large for the arena shows up.
int ensure_dev_ring(dm_engine *e)
{
    std::lock_guard<std::mutex> g(e->slab_mu);
    if (e->dev_ring) return DM_OK;
    cudaSetDevice(e->device);
    const uint64_t bytes = (uint64_t)e->slab_store.size() * e->cfg.slab_bytes;
    CU_TRY(cudaMalloc(&e->dev_ring, bytes));
    for (size_t i = 0; i < e->slab_store.size(); ++i) e->slab_store[i].dev = e->dev_ring + i * e->cfg.slab_bytes;
    return DM_OK;
}

std::shared_ptr<Stream> find_stream(dm_engine *e, uint64_t id)
{
    const int k = (int)(id % kStripes);
    std::lock_guard<std::mutex> g(e->stripe_mu[k]);
    auto it = e->streams[k].find(id);
    return it == e->streams[k].end() ? nullptr : it->second;
}

void drop_stream(dm_engine *e, const std::shared_ptr<Stream> &sp, bool release_slot)
{
    const int k = (int)(sp->id % kStripes);
    {
        std::lock_guard<std::mutex> g(e->stripe_mu[k]);
        if (e->streams[k].erase(sp->id)) e->n_streams--;
    }
    {
        std::lock_guard<std::mutex> g(e->mu);
        if (release_slot) e->free_slots.push_back(sp->slot);
        if (sp->has_expect) {
            auto it = e->inflight.find(sp->expect);
            if (it != e->inflight.end() && (it->second.expired() || it->second.lock() == sp)) e->inflight.erase(it);
        }
    }
}

int ensure_ingest_scratch(dm_engine *e, uint32_t n)
{
    if (n <= e->ing_cap) return DM_OK;
    const uint32_t cap = std::max<uint32_t>(n, 4096);
    if (e->ing_states) { cudaFree(e->ing_states); cudaFree(e->ing_digests); cudaFree(e->ing_jobs_d);
                         cudaFreeHost(e->ing_jobs_h); cudaFreeHost(e->ing_digests_h); e->ing_cap = 0; }
    CU_TRY(cudaMalloc(&e->ing_states, 32ull * cap));
    CU_TRY(cudaMalloc(&e->ing_digests, 32ull * cap));
    CU_TRY(cudaMalloc(&e->ing_jobs_d, sizeof(dm::HashJob) * (uint64_t)cap));
    CU_TRY(cudaHostAlloc(&e->ing_jobs_h, sizeof(dm::HashJob) * (uint64_t)cap, cudaHostAllocDefault));
    CU_TRY(cudaHostAlloc(&e->ing_digests_h, 32ull * cap, cudaHostAllocDefault));
    e->ing_cap = cap;
    return DM_OK;
}

}  // namespace

// ============================================================================
// C ABI
// ============================================================================
extern "C" {

uint32_t dm_abi_version(void) { return DM_ABI_VERSION; }

const char *dm_last_error(void) { return g_last_error.c_str(); }

const char *dm_strerror(int err)
{
    switch (err) {
    case DM_OK: return "ok";
    case DM_EINVAL: return "invalid argument";
    case DM_ENOMEM: return "out of HBM arena, pinned ring or stream slots";
    case DM_ENOENT: return "digest not in the content-addressed store";
    case DM_ECUDA: return "CUDA runtime error";
    case DM_ESTATE: return "call not valid in this stream state";
    case DM_EIO: return "disk tier I/O error";
    case DM_ENODEV: return "no usable CUDA device";
    case DM_ERANGE: return "offset beyond blob end";
    default: return "unknown error";
    }
}

int dm_device_count(void)
{
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess) { fail_cuda(err, "cudaGetDeviceCount"); return DM_ENODEV; }
    return n;
}

uint32_t dm_streams_per_warp(uint32_t n_resident) { return (uint32_t)dm::streams_per_warp_for(n_resident); }

uint32_t dm_shard_of(const uint8_t digest[32], uint32_t n_shards)
{
    if (!digest || n_shards <= 1) return 0;
    const uint32_t prefix = ((uint32_t)digest[0] << 8) | digest[1];
    return (uint32_t)(((uint64_t)prefix * n_shards) >> 16);
}

void dm_engine_destroy(dm_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    {
        std::lock_guard<std::mutex> g(e->work_mu);
        e->stop = true;
    }
    e->work_cv.notify_all();
    e->slab_cv.notify_all();
    if (e->pump.joinable()) e->pump.join();
    {
        std::lock_guard<std::mutex> g(e->spill_mu);
    }
    e->spill_cv.notify_all();
    for (auto &t : e->spillers) if (t.joinable()) t.join();
    cudaDeviceSynchronize();
    for (auto &m : e->readers) for (auto &kv : m) if (kv.second->fd >= 0) close(kv.second->fd);
    for (Cycle &c : e->cycles) {
        for (int i = 0; i < kCopyStreams; ++i) if (c.copy_ev[i]) cudaEventDestroy(c.copy_ev[i]);
        if (c.stream) cudaStreamDestroy(c.stream);
        if (c.k_start) cudaEventDestroy(c.k_start);
        if (c.k_end) cudaEventDestroy(c.k_end);
        if (c.h_jobs) cudaFreeHost(c.h_jobs);
        if (c.d_jobs) cudaFree(c.d_jobs);
    }
    for (SlabBatch &b : e->batches)
        for (int i = 0; i < kCopyStreams; ++i) if (b.ev[i]) cudaEventDestroy(b.ev[i]);
    for (Bounce &b : e->bounce_store) { if (b.host) cudaFreeHost(b.host); if (b.stream) cudaStreamDestroy(b.stream); }
    if (e->ing_states) { cudaFree(e->ing_states); cudaFree(e->ing_digests); cudaFree(e->ing_jobs_d);
                         cudaFreeHost(e->ing_jobs_h); cudaFreeHost(e->ing_digests_h); }
    if (e->ckpt_stream) cudaStreamDestroy(e->ckpt_stream);
    if (e->ckpt_pinned) cudaFreeHost(e->ckpt_pinned);
    if (e->ing_ev0) cudaEventDestroy(e->ing_ev0);
    if (e->ing_ev1) cudaEventDestroy(e->ing_ev1);
    if (e->d_states) cudaFree(e->d_states);
    if (e->h_digests) cudaFreeHost(e->h_digests);
    if (e->ring) cudaFreeHost(e->ring);
    if (e->dev_ring) cudaFree(e->dev_ring);
    if (e->arena_base) cudaFree(e->arena_base);
    for (int i = 0; i < kCopyStreams; ++i) if (e->copy_stream[i]) cudaStreamDestroy(e->copy_stream[i]);
    if (e->ingest_stream) cudaStreamDestroy(e->ingest_stream);
    if (e->util_stream) cudaStreamDestroy(e->util_stream);
    delete e;
}

int dm_engine_create(const dm_config *cfg, dm_engine **out)
{
    if (!cfg || !out || cfg->struct_size != sizeof(dm_config)) return fail(DM_EINVAL, "dm_config missing or wrong struct_size");
    *out = nullptr;
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0) {
        fail_cuda(err, "cudaGetDeviceCount");
        return DM_ENODEV;   // no CPU fallback on the hash path
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(DM_ENODEV, "device ordinal out of range");
    if (cfg->slab_bytes && (cfg->slab_bytes % 256)) return fail(DM_EINVAL, "slab_bytes must be a multiple of 256");

    dm_engine *e = new dm_engine();
    e->cfg = *cfg;
    e->device = cfg->device;
    if (cfg->cas_dir) e->cas_dir = cfg->cas_dir;
    e->cfg.cas_dir = nullptr;
    if (const char *v = getenv("DM_FORCE_SPW")) {
        const int f = atoi(v);
        if (f == 1 || f == 2 || f == 4 || f == 8 || f == 16 || f == 32) e->force_spw = f;
    }
    if (const char *v = getenv("DM_KERNEL_VARIANT")) {   // "wide,deep" variant numbers; experiments only
        int w = -1, d = -1;
        if (sscanf(v, "%d,%d", &w, &d) >= 1) {
            if (w >= 0 && w < 20) e->variant_wide = w;
            if (d >= 0 && d <= 4) e->variant_deep = d;   // 4 = short-chain round (deep and group kernels)
        }
    }
    if (!e->cfg.slab_bytes) e->cfg.slab_bytes = 1u << 20;
    if (!e->cfg.ring_bytes) e->cfg.ring_bytes = 256ull << 20;
    if (!e->cfg.max_streams) e->cfg.max_streams = 65536;
    if (e->cfg.ring_bytes < 4ull * e->cfg.slab_bytes) e->cfg.ring_bytes = 4ull * e->cfg.slab_bytes;

#define CU_INIT(expr)                                                                   \
    do {                                                                                \
        cudaError_t cu_err_ = (expr);                                                   \
        if (cu_err_ != cudaSuccess) { fail_cuda(cu_err_, #expr); dm_engine_destroy(e); return DM_ECUDA; } \
    } while (0)

    CU_INIT(cudaSetDevice(e->device));
    cudaDeviceProp prop;
    CU_INIT(cudaGetDeviceProperties(&prop, e->device));
    e->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) { fail(DM_ENODEV, "kernels are built for sm_100a only"); dm_engine_destroy(e); return DM_ENODEV; }

    for (int i = 0; i < kCopyStreams; ++i) CU_INIT(cudaStreamCreateWithFlags(&e->copy_stream[i], cudaStreamNonBlocking));
    CU_INIT(cudaStreamCreateWithFlags(&e->ingest_stream, cudaStreamNonBlocking));
    CU_INIT(cudaStreamCreateWithFlags(&e->util_stream, cudaStreamNonBlocking));
    CU_INIT(cudaStreamCreateWithFlags(&e->ckpt_stream, cudaStreamNonBlocking));
    CU_INIT(cudaHostAlloc(&e->ckpt_pinned, 64, cudaHostAllocDefault));
    CU_INIT(cudaEventCreate(&e->ing_ev0));
    CU_INIT(cudaEventCreate(&e->ing_ev1));

    if ((e->cfg.flags & DM_F_NO_HBM_CAS) && !e->cfg.hbm_cas_bytes) e->cfg.hbm_cas_bytes = 1u << 20;   // only dm_ingest_device would use it
    if (!e->cfg.hbm_cas_bytes) {
        size_t fr = 0, tot = 0;
        CU_INIT(cudaMemGetInfo(&fr, &tot));
        e->cfg.hbm_cas_bytes = fr / 2;
    }
    e->cfg.hbm_cas_bytes = round_up(e->cfg.hbm_cas_bytes, kAlign);
    CU_INIT(cudaMalloc(&e->arena_base, e->cfg.hbm_cas_bytes));
    e->arena.reset(e->cfg.hbm_cas_bytes);

    const uint64_t nslab = e->cfg.ring_bytes / e->cfg.slab_bytes;
    CU_INIT(cudaHostAlloc(&e->ring, nslab * e->cfg.slab_bytes, cudaHostAllocDefault));
    e->slab_store.resize(nslab);
    if (e->cfg.flags & DM_F_NO_HBM_CAS) CU_INIT(cudaMalloc(&e->dev_ring, nslab * e->cfg.slab_bytes));
    for (uint64_t i = 0; i < nslab; ++i) {
        e->slab_store[i].host = e->ring + i * e->cfg.slab_bytes;
        e->slab_store[i].dev = e->dev_ring ? e->dev_ring + i * e->cfg.slab_bytes : nullptr;
        e->slab_free.push_back(&e->slab_store[i]);
    }

    CU_INIT(cudaMalloc(&e->d_states, 32ull * e->cfg.max_streams));
    CU_INIT(cudaHostAlloc(&e->h_digests, 32ull * e->cfg.max_streams, cudaHostAllocMapped));
    CU_INIT(cudaHostGetDevicePointer((void **)&e->d_digests, e->h_digests, 0));
    e->free_slots.reserve(e->cfg.max_streams);
    for (uint32_t i = e->cfg.max_streams; i-- > 0;) e->free_slots.push_back(i);

    e->max_jobs = e->cfg.max_streams;
    for (Cycle &c : e->cycles) {
        for (int i = 0; i < kCopyStreams; ++i) CU_INIT(cudaEventCreateWithFlags(&c.copy_ev[i], cudaEventDisableTiming));
        CU_INIT(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        CU_INIT(cudaEventCreate(&c.k_start));
        CU_INIT(cudaEventCreate(&c.k_end));
        CU_INIT(cudaHostAlloc(&c.h_jobs, sizeof(dm::HashJob) * (uint64_t)e->max_jobs, cudaHostAllocDefault));
        CU_INIT(cudaMalloc(&c.d_jobs, sizeof(dm::HashJob) * (uint64_t)e->max_jobs));
    }
    for (SlabBatch &b : e->batches)
        for (int i = 0; i < kCopyStreams; ++i) CU_INIT(cudaEventCreateWithFlags(&b.ev[i], cudaEventDisableTiming));
    e->bounce_store.resize(kBounces);
    for (Bounce &b : e->bounce_store) {
        CU_INIT(cudaHostAlloc(&b.host, kBounceBytes, cudaHostAllocDefault));
        CU_INIT(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
        e->bounce_free.push_back(&b);
    }
#undef CU_INIT
    if (!e->cas_dir.empty()) mkdirs(e->cas_dir + "/blobs/sha256/x");
    e->pump = std::thread(pump_main, e);
    if (!e->cas_dir.empty())
        for (int i = 0; i < kSpillThreads; ++i) e->spillers.emplace_back(spill_main, e);
    *out = e;
    return DM_OK;
}

int dm_engine_stats(dm_engine *e, dm_stats *o)
{
    if (!e || !o) return fail(DM_EINVAL, "null argument");
    memset(o, 0, sizeof *o);
    o->bytes_ingested = e->st_ingested; o->bytes_hashed = e->st_hashed; o->bytes_served = e->st_served;
    o->blobs_committed = e->st_committed; o->blobs_mismatched = e->st_mismatch;
    o->kernel_launches = e->st_launches; o->launches_wide = e->st_wide; o->launches_deep = e->st_deep;
    { std::lock_guard<std::mutex> g(e->stat_mu); o->kernel_ms = e->st_kernel_ms; }
    o->h2d_bytes = e->st_h2d; o->d2h_bytes = e->st_d2h;
    { std::lock_guard<std::mutex> g(e->arena_mu); o->hbm_cas_used = e->arena.used(); o->hbm_cas_capacity = e->arena.capacity(); }
    o->open_streams = e->n_streams;
    o->ring_waits = e->st_ring_waits;
    o->launches_group = e->st_group;
    { std::lock_guard<std::mutex> g(e->slab_mu); o->ring_slabs_total = e->slab_store.size(); o->ring_slabs_free = e->slab_free.size(); }
    for (int k = 0; k < kStripes; ++k) { std::lock_guard<std::mutex> g(e->reader_mu[k]); o->open_readers += e->readers[k].size(); }
    { std::lock_guard<std::mutex> g(e->mu); o->free_stream_slots = e->free_slots.size(); }
    return DM_OK;
}

// ---- ingest ------------------------------------------------------------------

int dm_stream_open(dm_engine *e, const uint8_t expect[32], uint64_t size_hint, uint64_t *id)
{
    if (!e || !id) return fail(DM_EINVAL, "null argument");
    auto sp = std::make_shared<Stream>();
    if (expect) { sp->has_expect = true; memcpy(sp->expect.b, expect, 32); }
    {
        std::lock_guard<std::mutex> g(e->mu);
        if (e->free_slots.empty()) return fail(DM_ENOMEM, "max_streams reached");
        sp->slot = e->free_slots.back();
        e->free_slots.pop_back();
        sp->id = e->next_id++;
    }
    sp->verify_only = (e->cfg.flags & DM_F_NO_HBM_CAS) != 0;
    if (!sp->verify_only && size_hint > e->cfg.hbm_cas_bytes) {
        // can never be cached here: still verify it, through the device mirror of the ring
        int rc = ensure_dev_ring(e);
        if (rc != DM_OK) {
            std::lock_guard<std::mutex> g2(e->mu);
            e->free_slots.push_back(sp->slot);
            return rc;
        }
        sp->verify_only = true;
    }
    if (size_hint && !sp->verify_only) {
        cudaSetDevice(e->device);
        std::lock_guard<std::mutex> g(sp->mu);
        Extent x;
        if (!arena_alloc(e, size_hint, &x)) {
            std::lock_guard<std::mutex> g2(e->mu);
            e->free_slots.push_back(sp->slot);
            return fail(DM_ENOMEM, "HBM CAS arena exhausted");
        }
        sp->extents.push_back(x);
        sp->capacity = x.len;
    }
    sp->size_hint = size_hint;
    {
        const int k = (int)(sp->id % kStripes);
        std::lock_guard<std::mutex> g(e->stripe_mu[k]);
        e->streams[k][sp->id] = sp;
        e->n_streams++;
    }
    if (sp->has_expect && !sp->verify_only) {          // first opener wins; later duplicates are not followable
        std::lock_guard<std::mutex> g(e->mu);
        auto &slot = e->inflight[sp->expect];
        if (slot.expired()) slot = sp;
    }
    *id = sp->id;
    return DM_OK;
}

int dm_stream_write(dm_engine *e, uint64_t id, const void *buf, size_t len)
{
    if (!e || (!buf && len)) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    if (s->st != St::Open || s->window_out) return fail(DM_ESTATE, "stream not open for write");
    const uint8_t *p = static_cast<const uint8_t *>(buf);
    const uint32_t slab_bytes = e->cfg.slab_bytes;
    while (len) {
        if (!s->cur) {
            int rc = take_slab(e, s, g);
            if (rc != DM_OK) return rc;
        }
        const size_t n = std::min<size_t>(len, slab_bytes - s->cur_fill);
        if ((!s->islands.empty() || !s->parts.empty()) && range_taken(s, s->dma_issued + s->cur_fill, n, nullptr))
            return fail(DM_EINVAL, "write overlaps a range already received");
        memcpy(s->cur->host + s->cur_fill, p, n);
        s->cur_fill += (uint32_t)n; s->received += n; p += n; len -= n;
        if (s->cur_fill == slab_bytes) {
            int rc = submit_slab(e, sp);
            if (rc != DM_OK) return rc;
        }
    }
    return DM_OK;
}

int dm_stream_write_at(dm_engine *e, uint64_t id, uint64_t offset, const void *buf, size_t len)
{
    if (!e || (!buf && len)) return fail(DM_EINVAL, "null argument");
    if (offset + len < offset) return fail(DM_ERANGE, "offset + len overflows");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    if (s->st != St::Open || s->window_out) return fail(DM_ESTATE, "stream not open for write");
    if (offset == s->dma_issued + s->cur_fill + s->carry_fill && s->parts.empty() && s->islands.empty()) {
        g.unlock();
        return dm_stream_write(e, id, buf, len);          // plain sequential continuation
    }
    if (s->verify_only) return fail(DM_ESTATE, "out-of-order ranges need the HBM store (engine is verify-only)");
    const uint8_t *p = static_cast<const uint8_t *>(buf);
    const uint32_t slab_bytes = e->cfg.slab_bytes;
    while (len) {
        // the part that ends exactly here, or a new one
        size_t idx = s->parts.size();
        for (size_t i = 0; i < s->parts.size(); ++i)
            if (s->parts[i].base + s->parts[i].fill == offset) { idx = i; break; }
        if (idx == s->parts.size()) {
            if (offset == s->dma_issued + s->cur_fill) {
                // continues the contiguous run: use the sequential cursor
                if (!s->cur) {
                    int rc = take_slab(e, s, g);
                    if (rc != DM_OK) return rc;
                    continue;                               // state may have moved while unlocked
                }
                const size_t n = std::min<size_t>(len, slab_bytes - s->cur_fill);
                if (range_taken(s, offset, n, nullptr)) return fail(DM_EINVAL, "write overlaps a range already received");
                memcpy(s->cur->host + s->cur_fill, p, n);
                s->cur_fill += (uint32_t)n; s->received += n; p += n; len -= n; offset += n;
                if (s->cur_fill == slab_bytes) { int rc = submit_slab(e, sp); if (rc != DM_OK) return rc; }
                continue;
            }
            if (s->parts.size() >= 64) return fail(DM_ENOMEM, "too many concurrent range parts on one stream");
            if (range_taken(s, offset, 1, nullptr)) return fail(DM_EINVAL, "write overlaps a range already received");
            g.unlock();
            Slab *fresh = slab_get(e);
            g.lock();
            if (!fresh) return fail(DM_ESTATE, "engine stopping");
            if (s->st != St::Open) { slab_put(e, fresh); return fail(DM_ESTATE, "stream closed while waiting for the ring"); }
            s->parts.push_back({offset, fresh, 0});
            continue;                                       // re-find (the vector may have changed while unlocked)
        }
        Stream::Part &pt = s->parts[idx];
        const size_t n = std::min<size_t>(len, slab_bytes - pt.fill);
        if (range_taken(s, offset, n, &pt)) return fail(DM_EINVAL, "write overlaps a range already received");
        memcpy(pt.slab->host + pt.fill, p, n);
        pt.fill += (uint32_t)n; s->received += n; p += n; len -= n; offset += n;
        if (pt.fill == slab_bytes) { int rc = submit_part(e, sp, idx); if (rc != DM_OK) return rc; }
    }
    return DM_OK;
}

int dm_stream_checkpoint(dm_engine *e, uint64_t id, dm_checkpoint *out)
{
    if (!e || !out) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    if (s->st != St::Open || s->window_out) return fail(DM_ESTATE, "stream not open");
    if (s->cuda_failed) return fail(DM_ECUDA, "a CUDA copy or launch failed earlier on this stream: its state is not trusted");
    if (!s->verify_only || (s->cur_fill & 63) == 0) {
        // push out what is staged so the checkpoint covers every whole block received in order
        // (a verify-only stream hashes slab by slab, so only a block-aligned partial slab may go early)
        int rc = submit_slab(e, sp);
        if (rc != DM_OK) return rc;
    }
    // everything DMA'd so far in whole blocks must be hashed and no job may be running
    s->ckpt_waiter = true;
    s->cv.wait(g, [&] { return s->st != St::Open || (s->jobs_inflight == 0 && ((s->dma_issued - s->hash_issued) & ~63ull) == 0); });
    s->ckpt_waiter = false;
    if (s->st != St::Open) return fail(DM_ESTATE, "stream closed during checkpoint");
    if (s->cuda_failed) return fail(DM_ECUDA, "a CUDA copy or launch failed on this stream: its state is not trusted");
    memset(out, 0, sizeof *out);
    out->abi = DM_ABI_VERSION;
    out->bytes = s->hash_issued;
    if (s->hash_issued == 0) {
        static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au,
                                       0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        memcpy(out->h, iv, sizeof iv);
        return DM_OK;
    }
    cudaSetDevice(e->device);
    std::lock_guard<std::mutex> gc(e->ckpt_mu);
    CU_TRY(cudaMemcpyAsync(e->ckpt_pinned, e->d_states + 8ull * s->slot, 32, cudaMemcpyDeviceToHost, e->ckpt_stream));
    CU_TRY(cudaStreamSynchronize(e->ckpt_stream));
    memcpy(out->h, e->ckpt_pinned, 32);
    return DM_OK;
}

int dm_stream_resume(dm_engine *e, const dm_checkpoint *ck, const uint8_t expect[32], uint64_t size_hint, uint64_t *id)
{
    if (!e || !ck || !id) return fail(DM_EINVAL, "null argument");
    if (ck->abi != DM_ABI_VERSION || (ck->bytes & 63)) return fail(DM_EINVAL, "bad checkpoint");
    int rc = dm_stream_open(e, expect, size_hint, id);
    if (rc != DM_OK) return rc;
    auto sp = find_stream(e, *id);
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    s->resume_base = s->dma_issued = s->hash_issued = ck->bytes;
    if (ck->bytes && s->has_expect) {
        // not followable: the prefix may never be re-supplied, so there is nothing to serve from offset 0
        std::lock_guard<std::mutex> g2(e->mu);
        auto it = e->inflight.find(s->expect);
        if (it != e->inflight.end() && it->second.lock() == sp) e->inflight.erase(it);
    }
    if (ck->bytes) {
        // The state must be IN device memory before this returns: the stream's first job may launch at
        // once on another CUDA stream.  (A plain cudaMemcpy from pageable memory returns when the bytes
        // are staged, not when they have landed — found by tools/soak.py.)
        cudaSetDevice(e->device);
        std::lock_guard<std::mutex> gc(e->ckpt_mu);
        memcpy(e->ckpt_pinned, ck->h, 32);
        cudaError_t err = cudaMemcpyAsync(e->d_states + 8ull * s->slot, e->ckpt_pinned, 32, cudaMemcpyHostToDevice, e->ckpt_stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(e->ckpt_stream);
        if (err != cudaSuccess) {
            // nothing of this stream is in flight yet: give back its extent, state slot and id
            s->st = St::Aborted;
            g.unlock();
            free_extents(e, s->extents);
            drop_stream(e, sp, true);
            *id = 0;
            return fail_cuda(err, "cudaMemcpyAsync(checkpoint state)");
        }
    }
    return DM_OK;
}

int dm_stream_set_meta(dm_engine *e, uint64_t id, const char *key, const char *value)
{
    if (!e || !key || !value) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    std::lock_guard<std::mutex> g(sp->mu);
    if (sp->st != St::Open) return fail(DM_ESTATE, "stream not open");
    if (sp->meta.size() >= 64) return fail(DM_ENOMEM, "too many metadata entries");
    for (auto &kv : sp->meta) if (kv.first == key) { kv.second = value; return DM_OK; }
    sp->meta.emplace_back(key, value);
    return DM_OK;
}

int dm_stream_acquire(dm_engine *e, uint64_t id, void **ptr, size_t *cap)
{
    if (!e || !ptr || !cap) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    if (s->st != St::Open || s->window_out) return fail(DM_ESTATE, "stream not open or window outstanding");
    if (!s->cur) {
        int rc = take_slab(e, s, g);
        if (rc != DM_OK) return rc;
    }
    *ptr = s->cur->host + s->cur_fill;
    *cap = e->cfg.slab_bytes - s->cur_fill;
    s->window_out = true;
    return DM_OK;
}

int dm_stream_commit(dm_engine *e, uint64_t id, size_t len)
{
    if (!e) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    std::lock_guard<std::mutex> g(s->mu);
    if (!s->window_out) return fail(DM_ESTATE, "no window outstanding");
    if (len > e->cfg.slab_bytes - s->cur_fill) return fail(DM_EINVAL, "commit larger than the window");
    s->window_out = false;
    s->cur_fill += (uint32_t)len; s->received += len;
    if (s->cur_fill == e->cfg.slab_bytes) return submit_slab(e, sp);
    return DM_OK;
}

// Flush the partial slab and hand the stream to the pump for its final job.  Stream mutex held.
static int begin_finish(dm_engine *e, const std::shared_ptr<Stream> &sp, std::unique_lock<std::mutex> &g)
{
    Stream *s = sp.get();
    if (s->st == St::Finishing || s->st == St::Done) return DM_OK;
    if (s->st != St::Open || s->window_out) return fail(DM_ESTATE, "stream not open");
    if (s->cuda_failed) return fail(DM_ECUDA, "a CUDA copy or launch failed earlier on this stream (bytes may be missing): abort it");
    int rc = DM_OK;
    if (s->carry_fill && !s->cur) {                  // a recalled slab left a sub-block tail: it needs a slab to travel in
        rc = take_slab(e, s, g);
        if (rc != DM_OK) return rc;
        if (s->st == St::Finishing || s->st == St::Done) return DM_OK;      // someone else finished it while we waited
    }
    rc = submit_slab(e, sp);
    if (rc != DM_OK) return rc;
    while (!s->parts.empty()) {
        rc = submit_part(e, sp, s->parts.size() - 1);
        if (rc != DM_OK) return rc;
    }
    if (!s->islands.empty()) return fail(DM_ESTATE, "blob has holes: ranges missing before the last byte");
    s->st = St::Finishing;
    mark_dirty(e, sp, nullptr);
    return DM_OK;
}

int dm_stream_flush(dm_engine *e, uint64_t id)
{
    if (!e) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    std::unique_lock<std::mutex> g(sp->mu);
    return begin_finish(e, sp, g);
}

int dm_stream_finish(dm_engine *e, uint64_t id, uint8_t digest_out[32], int *matched)
{
    if (!e) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    std::shared_ptr<Blob> blob;
    bool failed = false;
    {
        std::unique_lock<std::mutex> g(s->mu);
        int rc = begin_finish(e, sp, g);
        if (rc != DM_OK) return rc;
        s->cv.wait(g, [&] { return s->st == St::Done; });
        if (digest_out) memcpy(digest_out, s->digest.b, 32);
        if (matched) *matched = s->matched;
        blob = s->blob;
        failed = s->cuda_failed;
    }
    drop_stream(e, sp, true);
    if (failed) return fail(DM_ECUDA, "a CUDA copy or launch failed while this stream was being hashed; nothing was cached");
    if (blob && (e->cfg.flags & DM_F_DISK_SYNC) && !e->cas_dir.empty()) {
        std::unique_lock<std::mutex> g(e->spill_mu);
        e->spill_done_cv.wait(g, [&] { std::lock_guard<std::mutex> g2(e->mu); return blob->spill_done; });
        std::lock_guard<std::mutex> g2(e->mu);
        if (!blob->on_disk) return fail(DM_EIO, "disk tier write failed");
    }
    return DM_OK;
}

int dm_stream_abort(dm_engine *e, uint64_t id)
{
    if (!e) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    bool free_now;
    {
        std::lock_guard<std::mutex> g(s->mu);
        if (s->st == St::Done || s->st == St::Aborted) return fail(DM_ESTATE, "stream already closed");
        if (s->cur) { slab_put(e, s->cur); s->cur = nullptr; s->cur_fill = 0; }
        s->carry_fill = 0;
        for (Stream::Part &pt : s->parts) slab_put(e, pt.slab);
        s->parts.clear();
        if (!s->staged.empty()) {            // their DMAs may be in flight: drain before the ring reuses them
            cudaSetDevice(e->device);
            cudaStreamSynchronize(e->copy_stream[s->id % kCopyStreams]);
            for (auto &ps : s->staged) slab_put(e, ps.first);
            s->staged.clear();
        }
        s->st = St::Aborted;
        free_now = s->jobs_inflight == 0;
    }
    if (free_now) {
        // A slab DMA into this extent may still be in flight; the range must not be handed to
        // another blob before it lands (the stale copy would overwrite the new owner's bytes).
        cudaSetDevice(e->device);
        cudaStreamSynchronize(e->copy_stream[s->id % kCopyStreams]);
        { std::unique_lock<std::mutex> g(s->mu); wait_follow_reads(s, g); }
        free_extents(e, s->extents);
    }
    drop_stream(e, sp, free_now);   // otherwise the pump releases slot + extents at reap (after the kernel,
                                    // which itself waited for every DMA enqueued before its launch)
    s->cv.notify_all();
    return DM_OK;
}

// ---- hit serving ---------------------------------------------------------------

int dm_cache_contains(dm_engine *e, const uint8_t digest[32], uint64_t *size)
{
    if (!e || !digest) return fail(DM_EINVAL, "null argument");
    Digest d;
    memcpy(d.b, digest, 32);
    {
        std::lock_guard<std::mutex> g(e->mu);
        auto it = e->blobs.find(d);
        if (it != e->blobs.end() && (it->second->in_hbm || it->second->on_disk)) {
            if (size) *size = it->second->size;
            return DM_OK;
        }
    }
    if (!e->cas_dir.empty()) {
        struct stat st;
        if (stat(blob_path(e, digest).c_str(), &st) == 0) { if (size) *size = (uint64_t)st.st_size; return DM_OK; }
    }
    return DM_ENOENT;
}

int dm_cache_open(dm_engine *e, const uint8_t digest[32], uint64_t *reader, uint64_t *size)
{
    if (!e || !digest || !reader) return fail(DM_EINVAL, "null argument");
    Digest d;
    memcpy(d.b, digest, 32);
    auto r = std::make_shared<Reader>();
    {
        std::lock_guard<std::mutex> g(e->mu);
        auto it = e->blobs.find(d);
        if (it != e->blobs.end() && it->second->in_hbm) {
            r->blob = it->second;
            r->blob->readers++;
            r->blob->tick = ++e->tick;
            r->size = r->blob->size;
        }
    }
    if (!r->blob) {
        if (e->cas_dir.empty()) return DM_ENOENT;
        r->fd = open(blob_path(e, digest).c_str(), O_RDONLY);
        if (r->fd < 0) return DM_ENOENT;
        struct stat st;
        fstat(r->fd, &st);
        r->size = (uint64_t)st.st_size;
        if (FILE *mf = fopen((blob_path(e, digest) + ".meta").c_str(), "r")) {
            char tmp[4096];
            size_t k;
            while ((k = fread(tmp, 1, sizeof tmp, mf)) > 0) r->disk_meta.append(tmp, k);
            fclose(mf);
        }
    }
    uint64_t id;
    {
        std::lock_guard<std::mutex> g(e->mu);
        id = e->next_id++;
    }
    {
        std::lock_guard<std::mutex> g(e->reader_mu[id % kStripes]);
        e->readers[id % kStripes][id] = r;
    }
    *reader = id;
    if (size) *size = r->size;
    return DM_OK;
}

static std::shared_ptr<Reader> find_reader(dm_engine *e, uint64_t id)
{
    std::lock_guard<std::mutex> g(e->reader_mu[id % kStripes]);
    auto it = e->readers[id % kStripes].find(id);
    return it == e->readers[id % kStripes].end() ? nullptr : it->second;
}

// Read from a body that is still arriving (request coalescing).  Blocks until bytes past `off` have
// been DMA'd, the body completes, or it fails.  Returns DM_OK with *nread set, a dm_err, or 1 when the
// body has completed and been published (the reader has been switched to the blob; caller continues).
static int follow_read(dm_engine *e, Reader *r, uint64_t off, void *buf, size_t len, size_t *nread)
{
    std::lock_guard<std::mutex> gr(r->mu);
    if (!r->follow) return 1;
    Stream *s = r->follow.get();
    std::vector<std::pair<uint8_t *, uint64_t>> segs;
    size_t n = 0;
    {
        std::unique_lock<std::mutex> g(s->mu);
        s->followers++;
        s->cv.wait(g, [&] {
            return s->st == St::Done || s->st == St::Aborted || (off < s->dma_issued && !s->completing) || len == 0;
        });
        s->followers--;
        if (s->st == St::Aborted) return fail(DM_ESTATE, "the upstream body this reader followed was aborted");
        if (s->st == St::Done) {
            std::shared_ptr<Blob> b = s->blob;
            g.unlock();
            if (!b) return fail(DM_ESTATE, "the upstream body this reader followed failed verification");
            std::lock_guard<std::mutex> g2(e->mu);
            if (!b->in_hbm) return fail(DM_ENOENT, "blob evicted before the follower switched over");
            b->readers++;
            r->blob = b;
            r->size = b->size;
            r->follow.reset();
            return 1;
        }
        if (len == 0) return DM_OK;
        n = (size_t)std::min<uint64_t>(len, s->dma_issued - off);
        for_segments(e, s->extents, off, n, [&](uint8_t *dev, uint64_t l) { segs.emplace_back(dev, l); });
        s->follow_reads++;                      // pins the extents until the copy-out below is done
    }
    struct Unpin {
        Stream *s;
        ~Unpin() { { std::lock_guard<std::mutex> g(s->mu); s->follow_reads--; } s->cv.notify_all(); }
    } unpin{s};
    // the bytes may still be in flight on the body's copy stream: order the read-back after them
    cudaSetDevice(e->device);
    Bounce *bn = bounce_get(e);
    cudaEvent_t ev;
    cudaError_t err = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventRecord(ev, e->copy_stream[s->id % kCopyStreams]);
    if (err == cudaSuccess) err = cudaStreamWaitEvent(bn->stream, ev, 0);
    uint8_t *out = static_cast<uint8_t *>(buf);
    size_t done = 0;
    for (auto &sg : segs) {
        uint64_t left = sg.second;
        uint8_t *dev = sg.first;
        while (left && err == cudaSuccess) {
            const size_t m = (size_t)std::min<uint64_t>(left, kBounceBytes);
            err = cudaMemcpyAsync(bn->host, dev, m, cudaMemcpyDeviceToHost, bn->stream);
            if (err == cudaSuccess) err = cudaStreamSynchronize(bn->stream);
            if (err == cudaSuccess) memcpy(out + done, bn->host, m);
            done += m; dev += m; left -= m;
        }
    }
    if (ev) cudaEventDestroy(ev);
    bounce_put(e, bn);
    if (err != cudaSuccess) return fail_cuda(err, "follow_read D2H");
    e->st_d2h += done; e->st_served += done;
    if (nread) *nread = done;
    return DM_OK;
}

// Start the D2H of [start, start + <=4 MiB) of the blob into a read-ahead window.
static cudaError_t window_fill(dm_engine *e, Reader *r, Window &w, uint64_t start)
{
    const uint64_t n = std::min<uint64_t>(kBounceBytes, r->size - start);
    uint8_t *dst = w.b->host;
    cudaError_t err = cudaSuccess;
    for_segments(e, r->blob->extents, start, n, [&](uint8_t *dev, uint64_t l) {
        if (err == cudaSuccess) err = cudaMemcpyAsync(dst, dev, l, cudaMemcpyDeviceToHost, w.b->stream);
        dst += l;
    });
    w.off = start; w.len = n; w.pending = true;
    e->st_d2h += n;
    return err;
}

int dm_cache_read(dm_engine *e, uint64_t reader, uint64_t off, void *buf, size_t len, size_t *nread)
{
    if (!e || (!buf && len)) return fail(DM_EINVAL, "null argument");
    std::shared_ptr<Reader> r = find_reader(e, reader);
    if (!r) return fail(DM_EINVAL, "unknown reader id");
    if (nread) *nread = 0;
    if (r->follow) {
        int rc = follow_read(e, r.get(), off, buf, len, nread);
        if (rc != 1) return rc;                 // 1: the body completed and was published: fall through to the blob
    }
    if (off > r->size) return fail(DM_ERANGE, "offset beyond blob end");
    len = (size_t)std::min<uint64_t>(len, r->size - off);
    if (len == 0) return DM_OK;
    uint8_t *out = static_cast<uint8_t *>(buf);
    if (!r->blob) {
        size_t got = 0;
        while (got < len) {
            ssize_t n = pread(r->fd, out + got, len - got, (off_t)(off + got));
            if (n < 0) { if (errno == EINTR) continue; return fail(DM_EIO, "pread failed"); }
            if (n == 0) break;
            got += (size_t)n;
        }
        if (nread) *nread = got;
        e->st_served += got;
        return DM_OK;
    }
    cudaSetDevice(e->device);
    std::lock_guard<std::mutex> g(r->mu);
    if (!r->tried_windows) {
        r->tried_windows = true;
        Bounce *a = bounce_try_get(e), *b = a ? bounce_try_get(e) : nullptr;
        if (a && b) { r->win[0].b = a; r->win[1].b = b; }
        else if (a) bounce_put(e, a);
    }
    size_t done = 0;
    int rc = DM_OK;
    if (r->win[0].b) {
        // HTTP bodies are read front to back in small pieces (io.Copy: 32 KiB): serve them from two
        // 4 MiB pinned windows, the next one filling by DMA while this one is copied out.
        while (done < len) {
            const uint64_t pos = off + done;
            Window *w = nullptr;
            for (Window &c : r->win) if (c.len && pos >= c.off && pos < c.off + c.len) w = &c;
            cudaError_t err = cudaSuccess;
            if (!w) {                                              // miss: restart the pipeline at pos
                for (Window &c : r->win) if (c.pending) { cudaStreamSynchronize(c.b->stream); c.pending = false; }
                err = window_fill(e, r.get(), r->win[0], pos);
                r->win[1].len = 0;
                if (err == cudaSuccess && pos + r->win[0].len < r->size)
                    err = window_fill(e, r.get(), r->win[1], pos + r->win[0].len);
                if (err != cudaSuccess) { rc = fail_cuda(err, "cudaMemcpyAsync(D2H)"); break; }
                w = &r->win[0];
            }
            if (w->pending) {
                err = cudaStreamSynchronize(w->b->stream);
                w->pending = false;
                if (err != cudaSuccess) { rc = fail_cuda(err, "cudaStreamSynchronize(D2H)"); break; }
            }
            const size_t n = (size_t)std::min<uint64_t>(len - done, w->off + w->len - pos);
            memcpy(out + done, w->b->host + (pos - w->off), n);
            done += n;
            if (pos + n == w->off + w->len) {                      // window drained: refill it behind the other one
                Window &other = (w == &r->win[0]) ? r->win[1] : r->win[0];
                const uint64_t next = other.len ? other.off + other.len : w->off + w->len;
                if (next < r->size && other.len && other.off == w->off + w->len) {
                    err = window_fill(e, r.get(), *w, next);
                    if (err != cudaSuccess) { rc = fail_cuda(err, "cudaMemcpyAsync(D2H)"); break; }
                } else w->len = 0;
            }
        }
    } else {
        Bounce *bn = bounce_get(e);                                // no windows left: one-shot staging
        while (done < len) {
            const size_t n = std::min(len - done, kBounceBytes);
            uint8_t *dst = bn->host;
            cudaError_t err = cudaSuccess;
            for_segments(e, r->blob->extents, off + done, n, [&](uint8_t *dev, uint64_t l) {
                if (err == cudaSuccess) err = cudaMemcpyAsync(dst, dev, l, cudaMemcpyDeviceToHost, bn->stream);
                dst += l;
            });
            if (err == cudaSuccess) err = cudaStreamSynchronize(bn->stream);
            if (err != cudaSuccess) { rc = fail_cuda(err, "cudaMemcpyAsync(D2H)"); break; }
            memcpy(out + done, bn->host, n);
            done += n;
            e->st_d2h += n;
        }
        bounce_put(e, bn);
    }
    e->st_served.fetch_add(done, std::memory_order_relaxed);
    if (nread) *nread = done;
    return rc;
}

int dm_cache_meta(dm_engine *e, uint64_t reader, char *buf, size_t cap, size_t *len)
{
    if (!e || !len || (!buf && cap)) return fail(DM_EINVAL, "null argument");
    std::shared_ptr<Reader> r = find_reader(e, reader);
    if (!r) return fail(DM_EINVAL, "unknown reader id");
    std::string j;
    if (r->blob) {
        std::lock_guard<std::mutex> g(e->mu);
        j = sidecar_json(*r->blob);
    } else j = r->disk_meta;
    *len = j.size();
    if (cap) {
        const size_t n = std::min(cap - 1, j.size());
        memcpy(buf, j.data(), n);
        buf[n] = 0;
    }
    return DM_OK;
}

int dm_cache_close(dm_engine *e, uint64_t reader)
{
    if (!e) return fail(DM_EINVAL, "null argument");
    std::shared_ptr<Reader> r;
    {
        std::lock_guard<std::mutex> g(e->reader_mu[reader % kStripes]);
        auto it = e->readers[reader % kStripes].find(reader);
        if (it == e->readers[reader % kStripes].end()) return fail(DM_EINVAL, "unknown reader id");
        r = it->second;
        e->readers[reader % kStripes].erase(it);
    }
    {
        std::lock_guard<std::mutex> g(r->mu);
        for (Window &w : r->win) {
            if (!w.b) continue;
            if (w.pending) cudaStreamSynchronize(w.b->stream);
            bounce_put(e, w.b);
            w.b = nullptr;
        }
    }
    if (r->blob) {
        std::lock_guard<std::mutex> g(e->mu);
        r->blob->readers--;
    }
    if (r->fd >= 0) close(r->fd);
    return DM_OK;
}

int dm_cache_follow(dm_engine *e, const uint8_t digest[32], uint64_t *reader, uint64_t *size_hint)
{
    if (!e || !digest || !reader) return fail(DM_EINVAL, "null argument");
    Digest d;
    memcpy(d.b, digest, 32);
    auto r = std::make_shared<Reader>();
    uint64_t id;
    {
        std::lock_guard<std::mutex> g(e->mu);
        auto it = e->inflight.find(d);
        if (it == e->inflight.end()) return DM_ENOENT;
        r->follow = it->second.lock();
        if (!r->follow) { e->inflight.erase(it); return DM_ENOENT; }
        id = e->next_id++;
    }
    r->size = r->follow->size_hint;
    {
        std::lock_guard<std::mutex> g(e->reader_mu[id % kStripes]);
        e->readers[id % kStripes][id] = r;
    }
    *reader = id;
    if (size_hint) *size_hint = r->size;
    return DM_OK;
}

int dm_cache_evict(dm_engine *e, const uint8_t digest[32])
{
    if (!e || !digest) return fail(DM_EINVAL, "null argument");
    Digest d;
    memcpy(d.b, digest, 32);
    std::shared_ptr<Blob> b;
    std::vector<Extent> ext;
    {
        std::lock_guard<std::mutex> g(e->mu);
        auto it = e->blobs.find(d);
        if (it == e->blobs.end() || !it->second->in_hbm) return DM_ENOENT;
        if (it->second->readers) return fail(DM_ESTATE, "blob has open readers");
        b = it->second;
        b->in_hbm = false;
        ext.swap(b->extents);                       // under the lock (see evict_for)
        if (!b->on_disk) e->blobs.erase(it);
    }
    free_extents(e, ext);
    return DM_OK;
}

int dm_cache_device_extents(dm_engine *e, uint64_t reader, void **dev_ptrs, uint64_t *lens, uint32_t max_ext)
{
    if (!e) return fail(DM_EINVAL, "null argument");
    std::shared_ptr<Reader> r = find_reader(e, reader);
    if (!r) return fail(DM_EINVAL, "unknown reader id");
    if (!r->blob) return fail(DM_ESTATE, "blob is on the disk tier only");
    const auto &ext = r->blob->extents;
    uint64_t left = r->blob->size;
    for (uint32_t i = 0; i < ext.size() && i < max_ext; ++i) {
        if (dev_ptrs) dev_ptrs[i] = e->arena_base + ext[i].off;
        if (lens) lens[i] = std::min(left, ext[i].len);
        left -= std::min(left, ext[i].len);
    }
    return (int)ext.size();
}

// ---- device-resident ingest -------------------------------------------------------

int dm_ingest_device(dm_engine *e, const void *dev_base, const uint64_t *offsets, const uint64_t *lengths,
                     uint32_t n, const uint8_t *expect, uint8_t *digests_out, uint8_t *matched_out,
                     uint32_t flags, double *kernel_ms)
{
    if (!e || ((!offsets || !lengths || !dev_base) && n)) return fail(DM_EINVAL, "null argument");
    if (kernel_ms) *kernel_ms = 0.0;
    if (n == 0) return DM_OK;
    if (((uintptr_t)dev_base) & 15) return fail(DM_EINVAL, "dev_base must be 16-byte aligned");
    for (uint32_t i = 0; i < n; ++i)
        if (offsets[i] & 15) return fail(DM_EINVAL, "offsets must be multiples of 16");
    cudaSetDevice(e->device);
    std::lock_guard<std::mutex> gi(e->ingest_mu);
    int rc = ensure_ingest_scratch(e, n);
    if (rc != DM_OK) return rc;
    const bool hash_only = (flags & DM_ING_HASH_ONLY) != 0;
    std::vector<Extent> ext(hash_only ? 0 : n, Extent{0, 0});
    auto cleanup = [&] {
        std::lock_guard<std::mutex> g(e->arena_mu);
        for (const Extent &x : ext) e->arena.release(x.off, x.len);
    };
    if ((flags & DM_ING_REPLACE) && expect && !hash_only) evict_many(e, expect, n);
    const uint8_t *base = static_cast<const uint8_t *>(dev_base);
    uint64_t total = 0;
    uint32_t allocated = 0;
    if (!hash_only) {                                   // fast path: all extents under one arena lock
        std::lock_guard<std::mutex> g(e->arena_mu);
        for (; allocated < n; ++allocated) {
            const uint64_t want = round_up(std::max<uint64_t>(lengths[allocated], 1), kAlign);
            uint64_t off;
            if (!e->arena.alloc(want, &off)) break;
            ext[allocated] = Extent{off, want};
        }
    }
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t len = lengths[i];
        dm::HashJob &jb = e->ing_jobs_h[i];
        jb.src = base + offsets[i]; jb.dst = nullptr; jb.nbytes = len; jb.total_len = len;
        jb.slot = i; jb.flags = dm::JOB_INIT | dm::JOB_FINAL; jb.one = 1; jb.pad_ = 0;
        if (!hash_only) {
            if (i >= allocated) {                       // arena full: evict LRU blobs one allocation at a time
                Extent x;
                if (!arena_alloc(e, len, &x)) { cleanup(); return fail(DM_ENOMEM, "HBM CAS arena exhausted"); }
                ext[i] = x;
            }
            jb.dst = e->arena_base + ext[i].off;
        }
        total += len;
    }
    int spw = e->force_spw ? e->force_spw : dm::streams_per_warp_for(n);
    if (flags & DM_ING_FORCE_WIDE) spw = 32;
    if (flags & DM_ING_FORCE_DEEP) spw = 1;
    if (flags & DM_ING_SPW_MASK) spw = 1 << (((flags & DM_ING_SPW_MASK) >> DM_ING_SPW_SHIFT) - 1);
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; ++i) order[i] = i;
    if (spw > 1) {
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return e->ing_jobs_h[a].nbytes > e->ing_jobs_h[b].nbytes; });
        std::vector<dm::HashJob> tmp(e->ing_jobs_h, e->ing_jobs_h + n);
        for (uint32_t i = 0; i < n; ++i) e->ing_jobs_h[i] = tmp[order[i]];   // slot keeps the caller's index
    }
    cudaStream_t st = e->ingest_stream;
    (void)cudaGetLastError();           // the caller's thread may carry a stale "not ready" from its own event polling
    cudaError_t err = cudaMemcpyAsync(e->ing_jobs_d, e->ing_jobs_h, sizeof(dm::HashJob) * (uint64_t)n, cudaMemcpyHostToDevice, st);
    if (err == cudaSuccess) err = cudaEventRecord(e->ing_ev0, st);
    if (err == cudaSuccess)
        err = spw == 1 ? dm::launch_sha256_deep(e->ing_jobs_d, n, e->ing_states, e->ing_digests, st, e->variant_deep)
            : spw == 32 ? dm::launch_sha256_wide(e->ing_jobs_d, n, e->ing_states, e->ing_digests, st, e->variant_wide)
                        : dm::launch_sha256_group(e->ing_jobs_d, n, e->ing_states, e->ing_digests, st, spw, e->variant_deep);
    if (err == cudaSuccess) err = cudaEventRecord(e->ing_ev1, st);
    if (err == cudaSuccess) err = cudaMemcpyAsync(e->ing_digests_h, e->ing_digests, 32ull * n, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) {
        cudaStreamSynchronize(st);      // whatever was enqueued before the failure still uses the job table and the extents
        cleanup();
        return fail_cuda(err, "dm_ingest_device launch");
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ing_ev0, e->ing_ev1);
    if (kernel_ms) *kernel_ms = ms;
    { std::lock_guard<std::mutex> g(e->stat_mu); e->st_kernel_ms += ms; }
    e->st_launches++; (spw == 1 ? e->st_deep : spw == 32 ? e->st_wide : e->st_group)++;
    e->st_hashed += total;
    std::vector<Verified> good;
    std::vector<Extent> bad;
    if (!hash_only) good.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
        Digest d;
        words_to_digest(e->ing_digests_h + 8ull * i, d.b);
        if (digests_out) memcpy(digests_out + 32ull * i, d.b, 32);
        const int ok = (!expect || memcmp(expect + 32ull * i, d.b, 32) == 0) ? 1 : 0;
        if (matched_out) matched_out[i] = (uint8_t)ok;
        if (hash_only) continue;
        if (ok) good.push_back(Verified{d, lengths[i], ext[i]});
        else { bad.push_back(ext[i]); e->st_mismatch++; }
    }
    if (!good.empty()) publish_many(e, good);
    if (!bad.empty()) {
        std::lock_guard<std::mutex> g(e->arena_mu);
        for (const Extent &x : bad) e->arena.release(x.off, x.len);
    }
    return DM_OK;
}

// ---- synthetic bytes ---------------------------------------------------------------

void dm_synth_fill_host(uint64_t seed, uint64_t blob, uint64_t byte_off, void *dst, size_t len)
{
    const uint64_t key = dm_blob_key(seed, blob);
    uint8_t *p = static_cast<uint8_t *>(dst);
    uint64_t j = byte_off;
    while (len && (j & 7)) { *p++ = (uint8_t)(dm_blob_word_k(key, j >> 3) >> (8 * (j & 7))); ++j; --len; }
    while (len >= 8) { const uint64_t w = dm_blob_word_k(key, j >> 3); memcpy(p, &w, 8); p += 8; j += 8; len -= 8; }
    while (len) { *p++ = (uint8_t)(dm_blob_word_k(key, j >> 3) >> (8 * (j & 7))); ++j; --len; }
}

int dm_synth_fill_device(dm_engine *e, uint64_t seed, uint64_t blob, uint64_t byte_off, void *dev_dst, size_t len)
{
    if (!e || (!dev_dst && len)) return fail(DM_EINVAL, "null argument");
    cudaSetDevice(e->device);
    CU_TRY(dm::launch_synth_fill(seed, blob, byte_off, dev_dst, len, e->util_stream));
    CU_TRY(cudaStreamSynchronize(e->util_stream));
    return DM_OK;
}

int dm_synth_fill_device_many(dm_engine *e, uint64_t seed, uint64_t first_blob, void *dev_base,
                              const uint64_t *offsets, const uint64_t *lengths, uint32_t n)
{
    if (!e || ((!offsets || !lengths || !dev_base) && n)) return fail(DM_EINVAL, "null argument");
    if (n == 0) return DM_OK;
    for (uint32_t i = 0; i < n; ++i) {
        if (offsets[i] & 15) return fail(DM_EINVAL, "offsets must be multiples of 16");
        if (i && offsets[i] < offsets[i - 1] + lengths[i - 1]) return fail(DM_EINVAL, "blobs must be ascending and disjoint");
    }
    cudaSetDevice(e->device);
    uint64_t *d_tab = nullptr;
    CU_TRY(cudaMalloc(&d_tab, 16ull * n));
    cudaError_t err = cudaMemcpyAsync(d_tab, offsets, 8ull * n, cudaMemcpyHostToDevice, e->util_stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(d_tab + n, lengths, 8ull * n, cudaMemcpyHostToDevice, e->util_stream);
    if (err == cudaSuccess)
        err = dm::launch_synth_fill_many(seed, first_blob, dev_base, d_tab, d_tab + n, n, offsets[0],
                                         offsets[n - 1] + lengths[n - 1] - offsets[0], e->util_stream);
    const cudaError_t sync = cudaStreamSynchronize(e->util_stream);     // also on failure: the table may still be in use
    if (err == cudaSuccess) err = sync;
    cudaFree(d_tab);
    if (err != cudaSuccess) return fail_cuda(err, "dm_synth_fill_device_many");
    return DM_OK;
}

}  // extern "C"
