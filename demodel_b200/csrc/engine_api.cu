// C-ABI, part 1: engine lifecycle and the ingest side (streams).  See engine_internal.hpp for the layout of the engine.
#include "engine_internal.hpp"

#include <dirent.h>
#include <sched.h>

namespace {

// DM_F_NUMA_LOCAL: the CPUs of the NUMA node the GPU hangs off (sysfs), or false when the topology is not
// visible (container without /sys/bus/pci, single-node box reporting -1).
bool numa_cpus_of_device(int device, cpu_set_t *set, int *node_out)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    for (char *c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[160];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return false;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return false;
    char list[4096] = {0};
    const size_t k = fread(list, 1, sizeof list - 1, f);
    fclose(f);
    list[k] = 0;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return false;
    CPU_ZERO(set);
    int n_set = 0;
    for (char *tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int lo = 0, hi = 0;
        const int m = sscanf(tok, "%d-%d", &lo, &hi);
        if (m < 1) continue;
        if (m == 1) hi = lo;
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) { CPU_SET(c, set); ++n_set; }
    }
    *node_out = node;
    return n_set > 0;
}

}  // namespace

extern "C" {

uint32_t dm_abi_version(void) { return DM_ABI_VERSION; }

const char *dm_last_error(void) { return g_last_error.c_str(); }

// Detail text by id instead of by thread.  cgo: a goroutine may run on a different OS thread by the time it
// asks for the text of a call that failed, so the thread-local dm_last_error() can return another
// connection's message; every failing stream / reader call also files its text under its id (0 for calls
// that have none), where any thread finds it.  e == NULL: the last dm_engine_create failure of the process.
static std::mutex g_create_err_mu;
static std::string g_create_err;
int dm_error_detail(dm_engine *e, uint64_t id, char *buf, size_t cap, size_t *len)
{
    if (!buf && cap) return fail(DM_EINVAL, "null argument");
    std::string t;
    if (!e) { std::lock_guard<std::mutex> g(g_create_err_mu); t = g_create_err; }
    else {
        std::lock_guard<std::mutex> g(e->err_mu);
        auto it = e->err_text.find(id);
        if (it != e->err_text.end()) t = it->second;
    }
    if (len) *len = t.size();
    if (cap) {
        const size_t n = std::min(cap - 1, t.size());
        memcpy(buf, t.data(), n);
        buf[n] = 0;
    }
    return DM_OK;
}
static int note_create_err(int rc)
{
    if (rc != DM_OK) { std::lock_guard<std::mutex> g(g_create_err_mu); g_create_err = g_last_error; }
    return rc;
}

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

uint32_t dm_default_kernel_variant(int which) { return (uint32_t)(which ? dm::kDefaultDeepVariant : dm::kDefaultWideVariant); }

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
    { std::lock_guard<std::mutex> g(e->done_mu); e->done_stop = true; }       // the pump has reaped everything: drain, then stop
    e->done_cv.notify_all();
    for (auto &t : e->completers) if (t.joinable()) t.join();
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
    if (e->ing_states) cudaFree(e->ing_states);
    if (e->ing_digests) cudaFree(e->ing_digests);
    if (e->ing_jobs_d) cudaFree(e->ing_jobs_d);
    if (e->ing_jobs_h) cudaFreeHost(e->ing_jobs_h);
    if (e->ing_digests_h) cudaFreeHost(e->ing_digests_h);
    if (e->alias_log) fclose(e->alias_log);
    if (e->ckpt_stream) cudaStreamDestroy(e->ckpt_stream);
    if (e->ckpt_pinned) cudaFreeHost(e->ckpt_pinned);
    if (e->ing_ev0) cudaEventDestroy(e->ing_ev0);
    if (e->ing_ev1) cudaEventDestroy(e->ing_ev1);
    if (e->ing_ev2) cudaEventDestroy(e->ing_ev2);
    for (int i = 0; i < kIngestMaxChunks; ++i) {
        if (e->ing_cev_k[i]) cudaEventDestroy(e->ing_cev_k[i]);
        if (e->ing_cev_done[i]) cudaEventDestroy(e->ing_cev_done[i]);
        if (e->ing_streams[i]) cudaStreamDestroy(e->ing_streams[i]);
    }
    if (e->pack_dev_base) cudaFree(e->pack_dev_base);
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

static int engine_create(const dm_config *cfg, dm_engine **out);
int dm_engine_create(const dm_config *cfg, dm_engine **out) { return note_create_err(engine_create(cfg, out)); }
static int engine_create(const dm_config *cfg, dm_engine **out)
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
            if (w >= 0 && w < 24) e->variant_wide = w;
            if (d >= 0 && d <= 10) e->variant_deep = d;   // 4..7 = short-chain rounds (deep and group kernels)
        }
    }
    // Streaming stores for the socket -> ring copy from 16 KiB pieces up (io.Copy moves 32 KiB); DM_NT_COPY_MIN=0
    // turns them off, any other value moves the threshold (A/B runs).
    e->nt_copy_min = 16384;
    if (const char *v = getenv("DM_NT_COPY_MIN")) e->nt_copy_min = (uint32_t)strtoul(v, nullptr, 10);
    if (const char *v = getenv("DM_INGEST_CHUNKS")) e->ingest_chunks = (uint32_t)std::min<long>(std::max<long>(atol(v), 0), kIngestMaxChunks);   // tuning / tests only
    if (const char *v = getenv("DM_SPLIT_MIN_BYTES")) e->split_min = strtoull(v, nullptr, 10);       // tuning / tests only
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
    // DM_F_NUMA_LOCAL: everything the engine pins (ring, bounce buffers, job tables) is first touched, and
    // every thread it starts (pump, spill) is created, while this thread is confined to the GPU's NUMA node;
    // the caller's affinity is restored before returning.  At 8 GPUs the ingest path is bound by host
    // memory traffic, and a ring on the wrong socket costs ~40 %.
    cpu_set_t saved_aff, node_aff;
    bool rebind = false;
    if ((e->cfg.flags & DM_F_NUMA_LOCAL) && sched_getaffinity(0, sizeof saved_aff, &saved_aff) == 0 &&
        numa_cpus_of_device(e->device, &node_aff, &e->numa_node) && sched_setaffinity(0, sizeof node_aff, &node_aff) == 0)
        rebind = true;
    else e->numa_node = -1;
    struct Restore {
        bool on; cpu_set_t *set;
        ~Restore() { if (on) sched_setaffinity(0, sizeof *set, set); }
    } restore{rebind, &saved_aff};
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
    CU_INIT(cudaEventCreate(&e->ing_ev2));
    for (int i = 0; i < kIngestMaxChunks; ++i) {
        CU_INIT(cudaStreamCreateWithFlags(&e->ing_streams[i], cudaStreamNonBlocking));
        CU_INIT(cudaEventCreate(&e->ing_cev_k[i]));
        CU_INIT(cudaEventCreateWithFlags(&e->ing_cev_done[i], cudaEventDisableTiming));
    }

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

    if (!(e->cfg.flags & DM_F_NO_HBM_CAS) && e->cfg.slab_bytes >= 1024) {      // staging for tiny-body packs
        e->tiny_max = std::min<uint32_t>(kTinyMax, e->cfg.slab_bytes / 4);
        CU_INIT(cudaMalloc(&e->pack_dev_base, (uint64_t)kPackDevBufs * e->cfg.slab_bytes));
        for (int i = 0; i < kPackDevBufs; ++i) e->pack_dev_free.push_back(e->pack_dev_base + (uint64_t)i * e->cfg.slab_bytes);
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
    if (!e->cas_dir.empty()) {
        mkdirs(e->cas_dir + "/blobs/sha256/x");
        mkdirs(e->cas_dir + "/partial/x");
        alias_load(e);
        if (DIR *dir = opendir((e->cas_dir + "/partial").c_str())) {      // downloads suspended by an earlier process
            while (struct dirent *de = readdir(dir)) {
                const size_t n = strlen(de->d_name);
                if (n == 69 && strcmp(de->d_name + 64, ".ckpt") == 0) e->n_suspended++;
            }
            closedir(dir);
        }
    }
    e->pump = std::thread(pump_main, e);
    for (int i = 0; i < kCompleters; ++i) e->completers.emplace_back(completer_main, e);
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
    { std::lock_guard<std::mutex> g(e->slot_mu); o->free_stream_slots = e->free_slots.size(); }
    o->numa_node = e->numa_node;
    { std::lock_guard<std::mutex> g(e->alias_mu); o->aliases = e->aliases.size(); }
    o->suspended = e->n_suspended;
    o->packed_bodies = e->st_packed; o->packs = e->st_packs;
    return DM_OK;
}

// ---- ingest ------------------------------------------------------------------

// `followable`: index the stream by its expected digest so that dm_cache_follow can attach to it.  A
// resumed stream is registered by dm_stream_resume itself, and only when it starts at byte 0: a follower
// reads from offset 0, and nothing before resume_base is ever written.
static int stream_open_impl(dm_engine *e, const uint8_t expect[32], uint64_t size_hint, uint64_t *id, bool followable)
{
    if (!e || !id) return fail(DM_EINVAL, "null argument");
    auto sp = std::make_shared<Stream>();
    if (expect) { sp->has_expect = true; memcpy(sp->expect.b, expect, 32); }
    {
        std::lock_guard<std::mutex> g(e->slot_mu);
        if (e->free_slots.empty()) return fail(DM_ENOMEM, "max_streams reached");
        sp->slot = e->free_slots.back();
        e->free_slots.pop_back();
    }
    sp->id = e->next_id.fetch_add(1);
    sp->verify_only = (e->cfg.flags & DM_F_NO_HBM_CAS) != 0;
    if (!sp->verify_only && size_hint > e->cfg.hbm_cas_bytes) {
        // can never be cached here: still verify it, through the device mirror of the ring
        int rc = ensure_dev_ring(e);
        if (rc != DM_OK) {
            std::lock_guard<std::mutex> g2(e->slot_mu);
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
            std::lock_guard<std::mutex> g2(e->slot_mu);
            e->free_slots.push_back(sp->slot);
            return fail(DM_ENOMEM, "HBM CAS arena exhausted");
        }
        sp->extents.push_back(x);
        sp->capacity = x.len;
    }
    sp->size_hint = size_hint;
    if (size_hint && size_hint <= e->tiny_max && !sp->verify_only) {       // announced tiny: no ring slab (see Stream::small)
        sp->small.reset(new uint8_t[size_hint]);
        sp->small_cap = (uint32_t)size_hint;
    }
    {
        const int k = (int)(sp->id % kStripes);
        std::lock_guard<std::mutex> g(e->stripe_mu[k]);
        e->streams[k][sp->id] = sp;
        e->n_streams++;
    }
    if (followable && sp->has_expect && !sp->verify_only) {          // first opener wins; later duplicates are not followable
        std::lock_guard<std::mutex> g(e->mu);
        auto &slot = e->inflight[sp->expect];
        if (slot.expired()) slot = sp;
    }
    *id = sp->id;
    return DM_OK;
}

int dm_stream_open(dm_engine *e, const uint8_t expect[32], uint64_t size_hint, uint64_t *id)
{
    return note_err(e, 0, stream_open_impl(e, expect, size_hint, id, true));
}

static const char *kLostText = "bytes this stream had accepted were dropped by an earlier failure (HBM arena full or a failed copy): abort it";

static int stream_write_impl(dm_engine *e, uint64_t id, const void *buf, size_t len)
{
    if (!e || (!buf && len)) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    if (s->st != St::Open || s->window_out) return fail(DM_ESTATE, "stream not open for write");
    if (s->lost != DM_OK) return fail(s->lost, kLostText);
    const uint8_t *p = static_cast<const uint8_t *>(buf);
    const uint32_t slab_bytes = e->cfg.slab_bytes;
    if (s->small_cap && !s->cur && s->small_fill + len <= s->small_cap) {     // announced-tiny body: private buffer, no slab
        if (len) memcpy(s->small.get() + s->small_fill, p, len);
        s->small_fill += (uint32_t)len; s->received += len;
        return DM_OK;
    }
    while (len) {
        if (s->lost != DM_OK) return fail(s->lost, kLostText);
        if (!s->cur) {
            int rc = take_slab(e, s, g);
            if (rc != DM_OK) return rc;
            continue;                                       // the pump may have marked the stream while we waited
        }
        const size_t n = std::min<size_t>(len, slab_bytes - s->cur_fill);
        if ((!s->islands.empty() || !s->parts.empty()) && range_taken(s, s->dma_issued + s->cur_fill, n, nullptr))
            return fail(DM_EINVAL, "write overlaps a range already received");
        ring_copy(e, s->cur->host + s->cur_fill, p, n);
        s->cur_fill += (uint32_t)n; s->received += n; p += n; len -= n;
        if (s->cur_fill == slab_bytes) {
            int rc = submit_slab(e, sp);
            if (rc != DM_OK) return rc;
        }
    }
    return DM_OK;
}

static int stream_write_at_impl(dm_engine *e, uint64_t id, uint64_t offset, const void *buf, size_t len)
{
    if (!e || (!buf && len)) return fail(DM_EINVAL, "null argument");
    if (offset + len < offset) return fail(DM_ERANGE, "offset + len overflows");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    if (s->st != St::Open || s->window_out) return fail(DM_ESTATE, "stream not open for write");
    if (s->lost != DM_OK) return fail(s->lost, kLostText);
    if (s->small_cap && !s->cur) {                       // range parts need real slabs: move the private buffer's bytes into one
        int rc = take_slab(e, s, g);
        if (rc != DM_OK) return rc;
    }
    if (offset < s->resume_base && offset + len > s->resume_base) {
        // A re-supplied prefix that runs across the resume point: the part below it is kept for caching
        // only (prefix_cover), the part above it is hashed.  Treat them as the two writes they are.
        const uint64_t below = s->resume_base - offset;
        g.unlock();
        int rc = stream_write_at_impl(e, id, offset, buf, (size_t)below);
        if (rc != DM_OK) return rc;
        return stream_write_at_impl(e, id, offset + below, static_cast<const uint8_t *>(buf) + below, len - (size_t)below);
    }
    if (offset == s->dma_issued + s->cur_fill + s->carry_fill && s->parts.empty() && s->islands.empty()) {
        g.unlock();
        return stream_write_impl(e, id, buf, len);          // plain sequential continuation
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
                ring_copy(e, s->cur->host + s->cur_fill, p, n);
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
            s->range_mode = true;
            continue;                                       // re-find (the vector may have changed while unlocked)
        }
        Stream::Part &pt = s->parts[idx];
        size_t n = std::min<size_t>(len, slab_bytes - pt.fill);
        // a part never straddles the resume point (submit_part files it as prefix cover OR as an island)
        const bool below = pt.base < s->resume_base;
        if (below) n = (size_t)std::min<uint64_t>(n, s->resume_base - offset);
        if (range_taken(s, offset, n, &pt)) return fail(DM_EINVAL, "write overlaps a range already received");
        ring_copy(e, pt.slab->host + pt.fill, p, n);
        pt.fill += (uint32_t)n; s->received += n; p += n; len -= n; offset += n;
        if (pt.fill == slab_bytes || (below && offset == s->resume_base)) { int rc = submit_part(e, sp, idx); if (rc != DM_OK) return rc; }
    }
    return DM_OK;
}

static int stream_checkpoint_impl(dm_engine *e, uint64_t id, dm_checkpoint *out)
{
    if (!e || !out) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    if (s->st != St::Open || s->window_out) return fail(DM_ESTATE, "stream not open");
    if (s->cuda_failed) return fail(DM_ECUDA, "a CUDA copy or launch failed earlier on this stream: its state is not trusted");
    if (s->lost != DM_OK) return fail(s->lost, kLostText);
    if (s->small_fill && !s->cur) {                      // bytes still in the private buffer of an announced-tiny body
        int rc = take_slab(e, s, g);
        if (rc != DM_OK) return rc;
        if (s->st != St::Open) return fail(DM_ESTATE, "stream closed during checkpoint");
    }
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
    if (s->lost != DM_OK) return fail(s->lost, kLostText);
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

static int stream_resume_impl(dm_engine *e, const dm_checkpoint *ck, const uint8_t expect[32], uint64_t size_hint, uint64_t *id);
int dm_stream_resume(dm_engine *e, const dm_checkpoint *ck, const uint8_t expect[32], uint64_t size_hint, uint64_t *id)
{
    return note_err(e, 0, stream_resume_impl(e, ck, expect, size_hint, id));
}
static int stream_resume_impl(dm_engine *e, const dm_checkpoint *ck, const uint8_t expect[32], uint64_t size_hint, uint64_t *id)
{
    if (!e || !ck || !id) return fail(DM_EINVAL, "null argument");
    if (ck->abi != DM_ABI_VERSION || (ck->bytes & 63)) return fail(DM_EINVAL, "bad checkpoint");
    // Not followable while it is being set up (a follower attaching now would read [0, ck->bytes) out of an
    // extent nothing has been written to), and never when it starts past byte 0: the prefix may not be
    // re-supplied at all, so there is nothing to serve from offset 0.
    int rc = stream_open_impl(e, expect, size_hint, id, false);
    if (rc != DM_OK) return rc;
    auto sp = find_stream(e, *id);
    if (!sp) return fail(DM_ESTATE, "stream closed while resuming");
    Stream *s = sp.get();
    std::unique_lock<std::mutex> g(s->mu);
    s->resume_base = s->dma_issued = s->hash_issued = s->landed = ck->bytes;
    if (ck->bytes == 0 && s->has_expect && !s->verify_only) {
        std::lock_guard<std::mutex> g2(e->mu);
        auto &slot = e->inflight[s->expect];
        if (slot.expired()) slot = sp;
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

// ---- checkpoints that survive a restart ------------------------------------
// <cas_dir>/partial/<hex of the expected digest>.ckpt   the record below
// <cas_dir>/partial/<hex>.part                           the first `bytes` bytes of the body
// The .ckpt is renamed into place last, so a record always has its bytes.

namespace {
struct SavedCkpt {
    char magic[8];          // "DMCKPT\0\2"
    uint32_t abi, reserved;
    uint64_t bytes;         // multiple of 64: hashed AND saved
    uint64_t size_hint;     // Content-Length the stream was opened with (0 = unknown)
    uint32_t h[8];          // SHA-256 chaining value after `bytes`
    uint8_t expect[32];
};
const char kCkptMagic[8] = {'D', 'M', 'C', 'K', 'P', 'T', 0, 2};
std::string partial_base(const dm_engine *e, const uint8_t d[32]) { return e->cas_dir + "/partial/" + hex_of(d, 32); }
}  // namespace

static int stream_abort_impl(dm_engine *e, uint64_t id);
static int stream_checkpoint_impl(dm_engine *e, uint64_t id, dm_checkpoint *out);
static int stream_resume_impl(dm_engine *e, const dm_checkpoint *ck, const uint8_t expect[32], uint64_t size_hint, uint64_t *id);

static int stream_suspend_impl(dm_engine *e, uint64_t id, uint64_t *resume_from)
{
    if (!e) return fail(DM_EINVAL, "null argument");
    if (resume_from) *resume_from = 0;
    if (e->cas_dir.empty()) return fail(DM_ESTATE, "dm_stream_suspend needs a disk tier (cas_dir)");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    Stream *s = sp.get();
    {
        std::lock_guard<std::mutex> g(s->mu);
        if (!s->has_expect) return fail(DM_ESTATE, "only a stream opened with an expected digest can be suspended (the digest names the saved files)");
        if (s->verify_only) return fail(DM_ESTATE, "a verify-only stream retains no bytes to save");
        const bool whole = s->resume_base == 0 ||
                           (s->prefix_cover.size() == 1 && s->prefix_cover.begin()->first == 0 &&
                            s->prefix_cover.begin()->second >= s->resume_base);
        if (!whole) return fail(DM_ESTATE, "the stream was resumed without its prefix: there is nothing whole to save");
    }
    dm_checkpoint ck;
    int rc = stream_checkpoint_impl(e, id, &ck);        // every whole block received in order is hashed now
    if (rc != DM_OK) return rc;
    std::vector<Extent> ext;
    uint64_t hint;
    {
        std::lock_guard<std::mutex> g(s->mu);
        if (s->st != St::Open) return fail(DM_ESTATE, "stream closed during suspend");
        ext = s->extents;
        hint = s->size_hint;
        s->follow_reads++;                               // pins the extents against abort / completion while they are copied out
    }
    cudaSetDevice(e->device);
    const std::string base = partial_base(e, s->expect.b);
    bool ok = cudaStreamSynchronize(e->copy_stream[s->id % kCopyStreams]) == cudaSuccess;     // the saved bytes have landed
    int fd = ok ? open((base + ".part.tmp").c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644) : -1;
    ok = ok && fd >= 0 && d2h_to_fd(e, ext, ck.bytes, fd);
    if (fd >= 0) ok = (close(fd) == 0) && ok;
    { std::lock_guard<std::mutex> g(s->mu); s->follow_reads--; }
    s->cv.notify_all();
    bool fresh = false;
    if (ok) {
        SavedCkpt rec;
        memset(&rec, 0, sizeof rec);
        memcpy(rec.magic, kCkptMagic, 8);
        rec.abi = DM_ABI_VERSION; rec.bytes = ck.bytes; rec.size_hint = hint;
        memcpy(rec.h, ck.h, sizeof rec.h);
        memcpy(rec.expect, s->expect.b, 32);
        struct stat st;
        fresh = stat((base + ".ckpt").c_str(), &st) != 0;
        unlink((base + ".ckpt").c_str());                // an older record must not pair with the new bytes
        ok = rename((base + ".part.tmp").c_str(), (base + ".part").c_str()) == 0;
        FILE *f = ok ? fopen((base + ".ckpt.tmp").c_str(), "wb") : nullptr;
        ok = f && fwrite(&rec, sizeof rec, 1, f) == 1 && fflush(f) == 0;
        if (f) fclose(f);
        ok = ok && rename((base + ".ckpt.tmp").c_str(), (base + ".ckpt").c_str()) == 0;
    }
    if (!ok) {
        unlink((base + ".part.tmp").c_str());
        unlink((base + ".ckpt.tmp").c_str());
        struct stat st;
        if (!fresh && stat((base + ".ckpt").c_str(), &st) != 0 && e->n_suspended) e->n_suspended--;    // the older record went with the attempt
        return fail(DM_EIO, "could not save the checkpoint under <cas_dir>/partial (the stream is still open)");
    }
    if (fresh) e->n_suspended++;
    stream_abort_impl(e, id);                            // releases the id, the state slot and the extent
    if (resume_from) *resume_from = ck.bytes;
    return DM_OK;
}

static int stream_resume_saved_impl(dm_engine *e, const uint8_t expect[32], uint64_t size_hint, uint64_t *id, uint64_t *resume_from)
{
    if (!e || !expect || !id) return fail(DM_EINVAL, "null argument");
    if (resume_from) *resume_from = 0;
    if (e->cas_dir.empty()) return DM_ENOENT;
    const std::string base = partial_base(e, expect);
    SavedCkpt rec;
    {
        FILE *f = fopen((base + ".ckpt").c_str(), "rb");
        if (!f) return DM_ENOENT;
        const bool got = fread(&rec, sizeof rec, 1, f) == 1;
        fclose(f);
        if (!got || memcmp(rec.magic, kCkptMagic, 8) != 0 || rec.abi != DM_ABI_VERSION || (rec.bytes & 63) ||
            memcmp(rec.expect, expect, 32) != 0) {
            unlink((base + ".ckpt").c_str()); unlink((base + ".part").c_str());
            if (e->n_suspended) e->n_suspended--;
            return fail(DM_EIO, "saved checkpoint is damaged or from another ABI: discarded");
        }
    }
    dm_checkpoint ck;
    memset(&ck, 0, sizeof ck);
    memcpy(ck.h, rec.h, sizeof ck.h);
    ck.bytes = rec.bytes; ck.abi = DM_ABI_VERSION;
    int rc = stream_resume_impl(e, &ck, expect, size_hint ? size_hint : rec.size_hint, id);
    if (rc != DM_OK) return rc;
    // Re-supply the saved prefix (cached, not re-hashed) through the ordinary range path: ring -> DMA -> extent.
    // A blob too large for the arena continues verify-only: it retains nothing, so there is no prefix to load.
    bool retains = true;
    if (auto sp = find_stream(e, *id)) { std::lock_guard<std::mutex> g(sp->mu); retains = !sp->verify_only; }
    int fd = rec.bytes && retains ? open((base + ".part").c_str(), O_RDONLY) : -1;
    bool ok = rec.bytes == 0 || !retains || fd >= 0;
    if (ok && rec.bytes && retains) {
        std::vector<uint8_t> buf(1u << 20);
        uint64_t off = 0;
        while (ok && off < rec.bytes) {
            const size_t want = (size_t)std::min<uint64_t>(buf.size(), rec.bytes - off);
            const ssize_t got = pread(fd, buf.data(), want, (off_t)off);
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) { ok = false; break; }
            rc = stream_write_at_impl(e, *id, off, buf.data(), (size_t)got);
            if (rc != DM_OK) ok = false;
            off += (uint64_t)got;
        }
    }
    if (fd >= 0) close(fd);
    if (!ok) {
        stream_abort_impl(e, *id);
        *id = 0;
        return rc != DM_OK ? rc : fail(DM_EIO, "could not read the saved bytes back from <cas_dir>/partial");
    }
    unlink((base + ".ckpt").c_str());                    // the stream owns the download again
    unlink((base + ".part").c_str());
    if (e->n_suspended) e->n_suspended--;
    if (resume_from) *resume_from = rec.bytes;
    return DM_OK;
}

static int stream_set_meta_impl(dm_engine *e, uint64_t id, const char *key, const char *value)
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

static int stream_acquire_impl(dm_engine *e, uint64_t id, void **ptr, size_t *cap)
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
    if (s->lost != DM_OK) return fail(s->lost, kLostText);
    *ptr = s->cur->host + s->cur_fill;
    *cap = e->cfg.slab_bytes - s->cur_fill;
    s->window_out = true;
    return DM_OK;
}

static int stream_commit_impl(dm_engine *e, uint64_t id, size_t len)
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
    if (s->lost != DM_OK) return fail(s->lost, kLostText);
    int rc = DM_OK;
    if (s->carry_fill && !s->cur) {                  // a recalled slab left a sub-block tail: it needs a slab to travel in
        rc = take_slab(e, s, g);
        if (rc != DM_OK) return rc;
        if (s->st == St::Finishing || s->st == St::Done) return DM_OK;      // someone else finished it while we waited
    }
    if (!pack_tiny_body(e, s)) {                       // tiny bodies share one DMA (struct Pack); everything else: its own
        if (s->small_fill && !s->cur) {                  // no pack to be had right now: the private buffer's bytes need a slab
            rc = take_slab(e, s, g);
            if (rc != DM_OK) return rc;
            if (s->st == St::Finishing || s->st == St::Done) return DM_OK;
        }
        rc = submit_slab(e, sp);
        if (rc != DM_OK) return rc;
    }
    while (!s->parts.empty()) {
        rc = submit_part(e, sp, s->parts.size() - 1);
        if (rc != DM_OK) return rc;
    }
    if (!s->islands.empty()) return fail(DM_ESTATE, "blob has holes: ranges missing before the last byte");
    if (s->lost != DM_OK) return fail(s->lost, kLostText);      // the pump may have recalled (and lost) a slab while take_slab waited
    s->st = St::Finishing;
    mark_dirty(e, sp, nullptr);
    return DM_OK;
}

static int stream_flush_impl(dm_engine *e, uint64_t id)
{
    if (!e) return fail(DM_EINVAL, "null argument");
    auto sp = find_stream(e, id);
    if (!sp) return fail(DM_EINVAL, "unknown stream id");
    std::unique_lock<std::mutex> g(sp->mu);
    return begin_finish(e, sp, g);
}

static int stream_finish_impl(dm_engine *e, uint64_t id, uint8_t digest_out[32], int *matched)
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
        s->cv.wait(g, [&] { return s->st == St::Done || s->st == St::Aborted; });
        // another thread closed the stream (BodyTee.Close from the client side) while this one waited for the
        // final hash: the abort owns the clean-up (id, state slot, extents); nothing was published
        if (s->st == St::Aborted) return fail(DM_ESTATE, "stream aborted while finishing");
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

static int stream_abort_impl(dm_engine *e, uint64_t id)
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
        s->small.reset(); s->small_fill = 0; s->small_cap = 0;
        for (Stream::Part &pt : s->parts) slab_put(e, pt.slab);
        s->parts.clear();
        if (!s->staged.empty()) {            // their DMAs may be in flight: drain before the ring reuses them
            cudaSetDevice(e->device);
            cudaStreamSynchronize(e->copy_stream[s->id % kCopyStreams]);
            for (auto &ps : s->staged) slab_return(e, ps.first);
            s->staged.clear();
        }
        s->st = St::Aborted;
        free_now = s->jobs_inflight == 0;
        if (free_now) pack_release_member(e, s);         // (a job already built keeps the pack until it is reaped)
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

// ---- exported wrappers: file the error text under the stream id (dm_error_detail) ----
int dm_stream_suspend(dm_engine *e, uint64_t id, uint64_t *resume_from) { return note_err(e, id, stream_suspend_impl(e, id, resume_from)); }
int dm_stream_resume_saved(dm_engine *e, const uint8_t expect[32], uint64_t size_hint, uint64_t *id, uint64_t *resume_from)
{
    return note_err(e, 0, stream_resume_saved_impl(e, expect, size_hint, id, resume_from));
}
int dm_stream_write(dm_engine *e, uint64_t id, const void *buf, size_t len) { return note_err(e, id, stream_write_impl(e, id, buf, len)); }
int dm_stream_write_at(dm_engine *e, uint64_t id, uint64_t offset, const void *buf, size_t len) { return note_err(e, id, stream_write_at_impl(e, id, offset, buf, len)); }
int dm_stream_checkpoint(dm_engine *e, uint64_t id, dm_checkpoint *out) { return note_err(e, id, stream_checkpoint_impl(e, id, out)); }
int dm_stream_set_meta(dm_engine *e, uint64_t id, const char *key, const char *value) { return note_err(e, id, stream_set_meta_impl(e, id, key, value)); }
int dm_stream_acquire(dm_engine *e, uint64_t id, void **ptr, size_t *cap) { return note_err(e, id, stream_acquire_impl(e, id, ptr, cap)); }
int dm_stream_commit(dm_engine *e, uint64_t id, size_t len) { return note_err(e, id, stream_commit_impl(e, id, len)); }
int dm_stream_flush(dm_engine *e, uint64_t id) { return note_err(e, id, stream_flush_impl(e, id)); }
int dm_stream_finish(dm_engine *e, uint64_t id, uint8_t digest_out[32], int *matched) { return note_err(e, id, stream_finish_impl(e, id, digest_out, matched)); }
int dm_stream_abort(dm_engine *e, uint64_t id) { return note_err(e, id, stream_abort_impl(e, id)); }

}  // extern "C"
