/*
 * demodel_b200.h — C-ABI of the B200 blob hash-and-cache engine.
 *
 * This is the drop-in boundary for demodel's cache path.  The reference
 * (moeru-ai/demodel @ fb8342aa) has no FFI, plugin or C interface of its own
 * (/root/reference/pkg/utils/fs.go:1-8 is all of pkg/), so
 * each entry point below cites the reference *hook* whose body it serves:
 *
 *   OnResponse hook  cmd/demodel/start.go:201-204   (today: Println, return resp)
 *       -> a Go io.ReadCloser wrapped around resp.Body calls
 *          dm_stream_open / dm_stream_write (or acquire+commit) /
 *          dm_stream_finish / dm_stream_abort        — "hash while copying
 *          into the cache".
 *   OnRequest hook   cmd/demodel/start.go:197-200   (today: Println, return req,nil)
 *       -> dm_cache_open / dm_cache_read / dm_cache_close build the
 *          *http.Response that short-circuits upstream — "serve a cache hit".
 *   start()          cmd/demodel/start.go:167-216
 *       -> dm_engine_create once per GPU before net.Listen (start.go:206),
 *          dm_engine_destroy on shutdown; dm_shard_of picks the engine.
 *
 * The cgo binding a maintainer would add is in INTEGRATION.md and go/.
 *
 * Conventions: plain C types only (no CUDA / torch types).  Every function is
 * thread-safe: net/http runs one goroutine per connection (start.go:210-215)
 * and cgo enters from many OS threads.  Functions return DM_OK (0) or a
 * negative dm_err; they never abort, throw or call exit.  There is NO CPU
 * fallback: without a usable CUDA device dm_engine_create fails with
 * DM_ENODEV / DM_ECUDA and nothing else can be called.
 */
#ifndef DEMODEL_B200_H
#define DEMODEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_ABI_VERSION 2u   /* 2: dm_error_detail, URL aliases, saved checkpoints, dm_gunzip, DM_F_NUMA_LOCAL; dm_stats grew */

typedef enum dm_err {
    DM_OK        = 0,
    DM_EINVAL    = -1,   /* bad argument / unknown id */
    DM_ENOMEM    = -2,   /* HBM arena, pinned ring or stream table exhausted */
    DM_ENOENT    = -3,   /* digest not in the content-addressed store */
    DM_ECUDA     = -4,   /* CUDA runtime error; dm_last_error() has the text */
    DM_ESTATE    = -5,   /* call not valid in the stream's current state */
    DM_EIO       = -6,   /* on-disk CAS tier I/O error */
    DM_ENODEV    = -7,   /* no CUDA device / device ordinal out of range */
    DM_ERANGE    = -8    /* offset beyond blob end */
} dm_err;

typedef struct dm_engine dm_engine;   /* opaque; owned by caller between create/destroy */

/* dm_config.flags */
#define DM_F_NO_HBM_CAS   0x1u  /* verify only: streams are hashed from a device mirror of the ring and
                                  * nothing is retained or published (also what a blob whose size_hint
                                  * exceeds hbm_cas_bytes gets automatically) */
#define DM_F_DISK_SYNC    0x2u  /* dm_stream_finish returns only after the disk tier holds the blob */
#define DM_F_NUMA_LOCAL   0x4u  /* pin the ring / bounce buffers and start the engine's threads on the NUMA node the
                                  * GPU hangs off (sysfs); silently skipped when the topology is not visible */

typedef struct dm_config {
    uint32_t struct_size;     /* = sizeof(dm_config); ABI guard */
    int32_t  device;          /* CUDA ordinal; one engine per GPU */
    uint64_t hbm_cas_bytes;   /* HBM arena for the content-addressed store; 0 = 1/2 of free HBM */
    uint64_t ring_bytes;      /* pinned host ring; 0 = 256 MiB */
    uint32_t slab_bytes;      /* ring slab = one H2D DMA; multiple of 256; 0 = 1 MiB */
    uint32_t max_streams;     /* concurrently open streams; 0 = 65536 */
    const char *cas_dir;      /* on-disk tier root, NULL = HBM tier only */
    uint32_t flags;
    uint32_t reserved;
} dm_config;

typedef struct dm_stats {
    uint64_t bytes_ingested;      /* through dm_stream_write/write_at/commit, counted per slab as it goes to the device */
    uint64_t bytes_hashed;        /* by the SHA-256 kernels */
    uint64_t bytes_served;        /* through dm_cache_read */
    uint64_t blobs_committed;
    uint64_t blobs_mismatched;
    uint64_t kernel_launches;     /* SHA-256 kernel launches (wide + deep) */
    uint64_t launches_wide;
    uint64_t launches_deep;
    double   kernel_ms;           /* CUDA-event time summed over those launches */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t hbm_cas_used;
    uint64_t hbm_cas_capacity;
    uint64_t open_streams;
    uint64_t ring_waits;          /* times a writer had to wait for a free ring slab (back-pressure) */
    uint64_t launches_group;      /* launches of the S-streams-per-warp kernel (counted in kernel_launches) */
    uint64_t ring_slabs_total;    /* leak accounting: every slab is either free, being filled, or in flight */
    uint64_t ring_slabs_free;
    uint64_t open_readers;
    uint64_t free_stream_slots;   /* == max_streams when no stream is open or draining */
    int64_t  numa_node;           /* DM_F_NUMA_LOCAL: the node the engine bound itself to; -1 = not bound */
    uint64_t aliases;             /* URL / ETag -> digest entries (dm_cache_alias_put) */
    uint64_t suspended;           /* interrupted downloads saved under <cas_dir>/partial (dm_stream_suspend) */
    uint64_t packed_bodies;       /* tiny bodies (<= 64 KiB, one piece) that shared a pack slab and its one H2D DMA */
    uint64_t packs;               /* ... and how many such packs went to the device */
} dm_stats;

/* ---- engine lifetime (start.go:167-216) -------------------------------- */
uint32_t    dm_abi_version(void);
int         dm_device_count(void);                       /* <0: dm_err */
int         dm_engine_create(const dm_config *cfg, dm_engine **out);
void        dm_engine_destroy(dm_engine *e);
int         dm_engine_stats(dm_engine *e, dm_stats *out);
const char *dm_strerror(int err);
const char *dm_last_error(void);                         /* thread-local detail text */
/* Detail text of the last failing call on stream / reader `id` (0: calls that take no id, such as
 * dm_stream_open, dm_cache_open, dm_ingest_device), retrievable from ANY thread - what a cgo caller
 * uses instead of dm_last_error(): a goroutine can change OS threads between the call that failed and
 * the one that asks why.  e == NULL: the last dm_engine_create failure in this process.  *len receives
 * the full length; up to cap-1 bytes + NUL are copied.  The engine remembers the last 4096 ids. */
int         dm_error_detail(dm_engine *e, uint64_t id, char *buf, size_t cap, size_t *len);

/* Kernel shape the engine picks for `n_resident` co-resident streams: streams per
 * warp, 1 = warp-per-stream (deep), 2..16 = group, 32 = lane-per-stream (wide).  Pure function. */
uint32_t    dm_streams_per_warp(uint32_t n_resident);

/* The kernel variant an engine uses unless DM_KERNEL_VARIANT overrides it: which = 0 the lane-per-stream (wide)
 * kernel's `fma + 4*style`, which = 1 the round form of the warp-per-stream / group kernels.  Reported next to
 * measurements (a profile taken with one variant says nothing about another).  Pure function. */
uint32_t    dm_default_kernel_variant(int which);

/* Digest-prefix sharding: which of n_shards engines owns a
 * blob.  Uses the leading 16 bits so any n_shards (not only powers of two)
 * splits the digest space evenly; for n = 2,4,8 it equals digest[0] >> (8-log2 n). */
uint32_t    dm_shard_of(const uint8_t digest[32], uint32_t n_shards);

/* ---- ingest: OnResponse body tee (start.go:201-204) -------------------- */
/* expect: the LFS oid / OCI layer digest known from the URL, or NULL.
 * size_hint: Content-Length, or 0 when unknown (chunked). */
int dm_stream_open(dm_engine *e, const uint8_t expect[32], uint64_t size_hint, uint64_t *id);
/* Copies buf before returning (the Go caller reuses its 32 KiB io.Copy
 * buffer on the next Read).  Blocks only for ring back-pressure. */
int dm_stream_write(dm_engine *e, uint64_t id, const void *buf, size_t len);
/* Range parts: bytes for blob offset `offset`, in any order (HF clients and
 * Ollama fetch one large blob as parallel `Range` requests).  The bytes land
 * at their place in the blob's HBM extent at once; the digest advances over the
 * contiguous prefix as it grows.  Ranges must not overlap; up to 64 parts may
 * be open on one stream.  dm_stream_finish fails with DM_ESTATE while holes remain. */
int dm_stream_write_at(dm_engine *e, uint64_t id, uint64_t offset, const void *buf, size_t len);

/* Mid-state checkpoint / resume of an interrupted download.
 * dm_stream_checkpoint waits until every whole block received in order is
 * hashed and returns (state words, byte count).  dm_stream_resume opens a new
 * stream that continues from such a checkpoint: bytes from `ck->bytes` on are
 * hashed; the prefix may be re-supplied with dm_stream_write_at purely for
 * caching — if it is not, the blob is verified but not published in the CAS. */
typedef struct dm_checkpoint {
    uint32_t h[8];      /* SHA-256 chaining value after `bytes` bytes */
    uint64_t bytes;     /* multiple of 64 */
    uint32_t abi;
    uint32_t reserved;
} dm_checkpoint;
int dm_stream_checkpoint(dm_engine *e, uint64_t id, dm_checkpoint *out);
int dm_stream_resume(dm_engine *e, const dm_checkpoint *ck, const uint8_t expect[32], uint64_t size_hint, uint64_t *id);

/* Interrupted downloads that survive a proxy restart (needs cas_dir).
 * dm_stream_suspend closes a stream opened WITH an expected digest and saves what it has: the SHA-256
 * mid-state after the last whole block received in order, and those bytes, under
 * <cas_dir>/partial/<hex>.ckpt / <hex>.part.  *resume_from = the byte count saved (multiple of 64): the
 * proxy re-requests upstream with `Range: bytes=<resume_from>-`.  The id is released.
 * dm_stream_resume_saved - in this process or after dm_engine_create over the same cas_dir - opens a stream
 * that continues from the saved state: the saved prefix is loaded back into the blob's HBM extent (so the
 * finished blob is cached whole) and only the bytes from *resume_from on are hashed.  DM_ENOENT when
 * nothing is saved for `expect`.  The saved files are removed once a resumed stream has taken them over. */
int dm_stream_suspend(dm_engine *e, uint64_t id, uint64_t *resume_from);
int dm_stream_resume_saved(dm_engine *e, const uint8_t expect[32], uint64_t size_hint, uint64_t *id, uint64_t *resume_from);

/* Response headers worth replaying on a hit (Content-Type, ETag, Last-Modified,
 * the request URL ...).  Stored with the blob — in the `.meta` sidecar on the
 * disk tier — and returned by dm_cache_meta.  At most 64 entries per stream. */
int dm_stream_set_meta(dm_engine *e, uint64_t id, const char *key, const char *value);

/* Zero-copy variant: borrow a window of the pinned ring, Read() into it,
 * then commit the bytes actually read.  At most one outstanding window per
 * stream; *cap >= 1 on success.
 * Lifetime: [ptr, ptr+cap) belongs to the caller from dm_stream_acquire until
 * dm_stream_commit RETURNS.  After the commit the engine may DMA the slab and
 * hand it to another stream at any moment, so anything else that reads the
 * window (the client-side write of the same bytes) must finish BEFORE the
 * commit: acquire, Read, Write to the client, commit. */
int dm_stream_acquire(dm_engine *e, uint64_t id, void **ptr, size_t *cap);
int dm_stream_commit(dm_engine *e, uint64_t id, size_t len);
/* Optional, non-blocking: upstream hit EOF, no more bytes will be written.
 * Starts the final hash so a later dm_stream_finish returns at once; lets one
 * OS thread multiplex many connections (goroutines on GOMAXPROCS threads). */
int dm_stream_flush(dm_engine *e, uint64_t id);
/* Blocks until every byte is hashed.  digest_out receives the SHA-256 of
 * the bytes written.  *matched = 1 if expect was NULL or equals the digest
 * (the blob is then published in the CAS under digest_out), 0 otherwise (the
 * bytes are discarded).  The id is released either way.
 * Errors: DM_ESTATE (holes left by range parts), DM_ENOMEM, DM_ECUDA before the
 * final hash was started leave the stream open - call dm_stream_abort.  DM_ECUDA
 * after it (a copy or launch failed while the body was being hashed): *matched = 0,
 * nothing was cached, the id is released. */
int dm_stream_finish(dm_engine *e, uint64_t id, uint8_t digest_out[32], int *matched);
/* Upstream error / client went away: drop the partial blob, release the id. */
int dm_stream_abort(dm_engine *e, uint64_t id);

/* ---- hit serving: OnRequest short-circuit (start.go:197-200) ----------- */
int dm_cache_contains(dm_engine *e, const uint8_t digest[32], uint64_t *size);  /* DM_ENOENT on miss */
int dm_cache_open(dm_engine *e, const uint8_t digest[32], uint64_t *reader, uint64_t *size);
int dm_cache_read(dm_engine *e, uint64_t reader, uint64_t off, void *buf, size_t len, size_t *nread);
/* The blob's sidecar as JSON text: {"digest":"sha256:..","size":N,"encoding":"identity",
 * ..,"headers":{..}}.  *len receives the full length; up to cap-1 bytes + NUL are copied. */
int dm_cache_meta(dm_engine *e, uint64_t reader, char *buf, size_t cap, size_t *len);
int dm_cache_close(dm_engine *e, uint64_t reader);
int dm_cache_evict(dm_engine *e, const uint8_t digest[32]);   /* HBM tier only; disk copy stays */
/* Request coalescing: the blob is not cached yet but another connection is
 * ingesting it right now (a stream opened with this expected digest).  The
 * reader follows that body out of HBM: dm_cache_read blocks until bytes past
 * `off` have arrived, and once the body completes and verifies the reader
 * becomes an ordinary cache reader.  If the body is aborted or fails
 * verification, reads fail with DM_ESTATE.  DM_ENOENT: nothing in flight.
 * *size_hint = the Content-Length the ingesting stream was opened with (0 = unknown). */
int dm_cache_follow(dm_engine *e, const uint8_t digest[32], uint64_t *reader, uint64_t *size_hint);

/* URL -> digest index for the OnRequest hit check (start.go:197-200 hands the hook a request, not a
 * digest).  OCI blob URLs carry `sha256:<hex>`; HuggingFace `resolve/<rev>/<file>` URLs (the first clients
 * /root/reference/README.md:16-21 lists) do not.  The tee records URL (or ETag / LFS oid string) -> digest
 * when a body verifies; a later request for the same key is answered from the CAS.  `key` is any
 * NUL-terminated string up to 4096 bytes.  With cas_dir the index is an append-only log
 * (<cas_dir>/aliases.log) replayed by dm_engine_create; a put for an existing key replaces it.
 * dm_cache_alias_get: DM_ENOENT when the key is unknown (the digest may since have been evicted from both
 * tiers: dm_cache_open decides). */
int dm_cache_alias_put(dm_engine *e, const char *key, const uint8_t digest[32]);
int dm_cache_alias_get(dm_engine *e, const char *key, uint8_t digest_out[32]);

/* ---- device-resident ingest -------------------------------------------- */
/* Hash-and-cache n blobs whose bytes are ALREADY in this GPU's HBM (landed
 * by GPUDirect, produced on device, or the bench's "inputs resident" leg).
 * Blob i occupies [dev_base + offsets[i], dev_base + offsets[i] + lengths[i]);
 * every offsets[i] must be a multiple of 16 and the blobs must not overlap.
 * expect (n*32 bytes) may be NULL.
 * One fused kernel pass reads each byte once, advances its blob's SHA-256
 * state and writes the byte into the blob's CAS extent.  digests_out gets
 * n*32 bytes; matched_out (n bytes, optional) the per-blob verdicts.
 * kernel_ms (optional) = CUDA-event time of the hash launches on the
 * engine's own stream(s): a batch of 32 768 or more small blobs goes out as
 * up to 8 launches on as many streams, so that extents, verdicts and index
 * updates of one part overlap the kernel of another; kernel_ms then runs
 * from the first launch to the end of the last.  flags: DM_ING_* below. */
#define DM_ING_HASH_ONLY   0x1u   /* do not copy into the CAS (1 B/B of traffic instead of 2) */
#define DM_ING_REPLACE     0x2u   /* re-ingest: a cached blob with an expected digest is taken out of sight for the
                                   * call and - when it has no reader - its extent is rewritten in place */
#define DM_ING_FORCE_WIDE  0x4u   /* kernel selection override: lane-per-stream */
#define DM_ING_FORCE_DEEP  0x8u   /* kernel selection override: warp-per-stream */
#define DM_ING_SPW_SHIFT   8      /* kernel selection override: (log2(streams per warp) + 1) << 8, */
#define DM_ING_SPW_MASK    0x700u /*   i.e. 1..6 for 1, 2, 4, 8, 16, 32 streams per warp          */
int dm_ingest_device(dm_engine *e, const void *dev_base, const uint64_t *offsets,
                     const uint64_t *lengths, uint32_t n, const uint8_t *expect, uint8_t *digests_out, uint8_t *matched_out,
                     uint32_t flags, double *kernel_ms);

/* Device pointer + extent walk of a cached blob, for consumers that want the
 * bytes in HBM (e.g. a weight loader on the same GPU).  Returns the number
 * of extents; fills up to max_ext (dev_ptr, len) pairs. */
int dm_cache_device_extents(dm_engine *e, uint64_t reader, void **dev_ptrs, uint64_t *lens, uint32_t max_ext);

/* ---- manifest-aware prefetch --------------------------- */
/* An OCI / Ollama image manifest (shape: the reference's cached fixture,
 * CONTRIBUTING.md:128-153) lists every blob with digest and size before the
 * bodies are requested. */
typedef struct dm_layer {
    uint8_t  digest[32];
    uint64_t size;
    char     media_type[96];   /* NUL-terminated, truncated if longer */
} dm_layer;
/* Collects every descriptor (object with a "sha256:" digest and an integer
 * size) in document order: config first, then layers.  *n_layers receives the
 * total found even if it exceeds max_layers.  DM_EINVAL on malformed JSON. */
int dm_manifest_parse(const char *json, size_t len, dm_layer *out, uint32_t max_layers, uint32_t *n_layers);
/* Pre-open one stream per layer that is not already cached, carrying its
 * expected digest and size (so its extent is reserved and the body is verified
 * as it arrives).  ids[i] = 0 for layers that are cache hits or duplicates. */
int dm_manifest_prefetch(dm_engine *e, const dm_layer *layers, uint32_t n, uint64_t *ids);
/* Registries answer manifest requests with `Content-Encoding: gzip` when the client allows it - the one
 * cached body the reference documents is such a response (CONTRIBUTING.md:76-99,116).  Inflates a gzip
 * (RFC 1952) or raw-deflate-in-zlib (RFC 1950) body of `len` bytes into dst[cap]; *out_len = inflated size.
 * DM_EINVAL: malformed stream or CRC / length mismatch; DM_ENOMEM: cap too small (*out_len = needed size
 * when the trailer states it).  Pure host code; manifests are a few KiB. */
int dm_gunzip(const void *src, size_t len, void *dst, size_t cap, size_t *out_len);

/* ---- synthetic blob bytes ------------------------------ */
/* Counter-based generator: byte j of blob `blob` under `seed` is a pure
 * function of (seed, blob, j); host and device versions agree bit for bit. */
void dm_synth_fill_host(uint64_t seed, uint64_t blob, uint64_t byte_off, void *dst, size_t len);
int  dm_synth_fill_device(dm_engine *e, uint64_t seed, uint64_t blob, uint64_t byte_off,
                          void *dev_dst, size_t len);
/* Fill n blobs laid out by offsets[]/lengths[] (as in dm_ingest_device, offsets
 * ascending) in one launch; blob i gets generator index first_blob + i. */
int  dm_synth_fill_device_many(dm_engine *e, uint64_t seed, uint64_t first_blob, void *dev_base,
                               const uint64_t *offsets, const uint64_t *lengths, uint32_t n);

/* ---- proxy-side driver (the host half of the path, in-process) --------- */
/* What the Go side does per connection, restated in C++ threads so the
 * end-to-end path can be driven here, where no Go toolchain exists:
 * `nthreads` connection workers each pull whole blobs from a shared queue,
 * read them in `chunk`-byte pieces from caller-owned HOST memory
 * [host_base + offsets[i], host_base + offsets[i+1]) and push them through
 * dm_stream_open/write/finish.  zero_copy selects how: 0 copy + non-blocking
 * EOF (flush), 1 acquire/commit with the "socket read" landing directly in the
 * ring, 2 / 3 the blocking io.ReadCloser forms of 0 / 1 (one body per thread at
 * a time), 4 as 2 but the client disconnects half way (Close before EOF:
 * matched_out[i] = 2, nothing is published).  Streams are interleaved
 * `concurrency` at a time the way concurrent goroutines would be.
 * Returns wall seconds in *seconds. */
int dm_proxy_drive(dm_engine *e, const void *host_base, const uint64_t *offsets, uint32_t n,
                   const uint8_t *expect, size_t chunk, uint32_t concurrency, int nthreads,
                   int zero_copy, uint8_t *digests_out, uint8_t *matched_out, double *seconds);
/* Serve n cached blobs back out into caller-owned host memory in `chunk`
 * pieces (the cache-hit path), `nthreads` readers. */
int dm_proxy_serve(dm_engine *e, const uint8_t *digests, uint32_t n, void *host_base,
                   const uint64_t *offsets, size_t chunk, int nthreads, double *seconds);

/* Single-connection twins of the two hooks for callers that work with URLs (demodel_b200/csrc/proxy_hooks.hpp:
 * BodyTee::SetURL, HitReader(url), ManifestTee).
 * dm_proxy_fetch: the OnResponse tee for one body fetched under `url`; expect may be NULL (HuggingFace resolve/
 * URL: digest unknown until hashed).  A verified body is entered in the alias index under the URL.
 * dm_proxy_request: the OnRequest check - a digest in the URL (OCI `sha256:<hex>`) or the alias index.
 * DM_ENOENT = miss; DM_OK = *reader is an open cache reader over *size bytes.
 * dm_proxy_manifest: the OnResponse tee for a manifest response, Content-Encoding "gzip" or identity (NULL):
 * passes the body through, inflates, parses, prefetches.  Up to max_layers entries are returned. */
int dm_proxy_fetch(dm_engine *e, const char *url, const void *body, uint64_t len, const uint8_t *expect,
                   size_t chunk, uint8_t digest_out[32], int *matched_out);
int dm_proxy_request(dm_engine *e, const char *url, uint64_t *reader, uint64_t *size);
int dm_proxy_manifest(dm_engine *e, const void *body, uint64_t len, const char *content_encoding, size_t chunk,
                      dm_layer *layers_out, uint64_t *ids_out, uint32_t max_layers, uint32_t *n_layers);

#ifdef __cplusplus
}
#endif
#endif /* DEMODEL_B200_H */
