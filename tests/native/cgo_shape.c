/* Enters the C-ABI the way cgo does - from plain C (this file is C99: the header must be usable without C++),
 * on SHORT-LIVED FOREIGN THREADS that the library never saw before and never sees again, with consecutive calls
 * on one stream / reader coming from DIFFERENT threads (a goroutine migrates between OS threads between cgo
 * calls), and with the error text of a failed call fetched on yet another thread (dm_error_detail, not the
 * thread-local dm_last_error).  The library may not keep anything a call needs in thread-local state.
 *
 * Linked against the real libdemodel_b200.so under `-m gpu` (tests/test_cgo_shape.py) and against the
 * fake-runtime build on the CPU box.  Digests are checked against FIPS 180-4 known answers ("abc", one million
 * 'a') and for self-consistency (expect = the digest a first pass produced; served bytes = ingested bytes).
 *
 * usage: cgo_shape [cas_dir]        exit 0 = ok
 */
#define _POSIX_C_SOURCE 200809L
#include "../../include/demodel_b200.h"

#include <pthread.h>
#include <time.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static dm_engine *E;
static int failures;

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); failures++; } } while (0)

/* one cgo call = one closure run on a brand-new thread that exits afterwards */
typedef struct call { void (*fn)(struct call *); int rc; uint64_t id, u64; const void *in; void *out; size_t len, got; const uint8_t *digest;
                      uint8_t dig[32]; int matched; char text[256]; const char *str; dm_checkpoint ck; } call;
static void *tramp(void *p) { call *c = (call *)p; c->fn(c); return NULL; }
static void on_new_thread(call *c)
{
    pthread_t t;
    if (pthread_create(&t, NULL, tramp, c) != 0) { fprintf(stderr, "pthread_create failed\n"); exit(2); }
    pthread_join(t, NULL);
}

static void c_open(call *c) { c->rc = dm_stream_open(E, c->digest, c->u64, &c->id); }
static void c_write(call *c) { c->rc = dm_stream_write(E, c->id, c->in, c->len); }
static void c_write_at(call *c) { c->rc = dm_stream_write_at(E, c->id, c->u64, c->in, c->len); }
static void c_flush(call *c) { c->rc = dm_stream_flush(E, c->id); }
static void c_finish(call *c) { c->rc = dm_stream_finish(E, c->id, c->dig, &c->matched); }
static void c_abort(call *c) { c->rc = dm_stream_abort(E, c->id); }
static void c_meta(call *c) { c->rc = dm_stream_set_meta(E, c->id, "ETag", c->str); }
static void c_acquire(call *c) { c->rc = dm_stream_acquire(E, c->id, &c->out, &c->len); }
static void c_commit(call *c) { c->rc = dm_stream_commit(E, c->id, c->len); }
static void c_ckpt(call *c) { c->rc = dm_stream_checkpoint(E, c->id, &c->ck); }
static void c_resume(call *c) { c->rc = dm_stream_resume(E, &c->ck, c->digest, c->u64, &c->id); }
static void c_suspend(call *c) { c->rc = dm_stream_suspend(E, c->id, &c->u64); }
static void c_resume_saved(call *c) { c->rc = dm_stream_resume_saved(E, c->digest, 0, &c->id, &c->u64); }
static void c_copen(call *c) { c->rc = dm_cache_open(E, c->digest, &c->id, &c->u64); }
static void c_cread(call *c) { c->rc = dm_cache_read(E, c->id, c->u64, c->out, c->len, &c->got); }
static void c_cmeta(call *c) { c->rc = dm_cache_meta(E, c->id, c->text, sizeof c->text, &c->got); }
static void c_cclose(call *c) { c->rc = dm_cache_close(E, c->id); }
static void c_cevict(call *c)
{   /* with a disk tier the spill thread may hold the blob for a moment (DM_ESTATE): retry briefly */
    for (int i = 0; i < 2000; ++i) {
        c->rc = dm_cache_evict(E, c->digest);
        if (c->rc != DM_ESTATE) break;
        struct timespec ts = {0, 1000000};
        nanosleep(&ts, NULL);
    }
}
static void c_contains(call *c) { c->rc = dm_cache_contains(E, c->digest, &c->u64); }
static void c_follow(call *c) { c->rc = dm_cache_follow(E, c->digest, &c->id, &c->u64); }
static void c_alias_put(call *c) { c->rc = dm_cache_alias_put(E, c->str, c->digest); }
static void c_alias_get(call *c) { c->rc = dm_cache_alias_get(E, c->str, c->dig); }
static void c_request(call *c) { c->rc = dm_proxy_request(E, c->str, &c->id, &c->u64); }
static void c_detail(call *c) { c->rc = dm_error_detail(E, c->id, c->text, sizeof c->text, &c->got); }
static void c_stats(call *c) { dm_stats st; c->rc = dm_engine_stats(E, &st); c->u64 = st.open_streams + st.open_readers; c->got = (size_t)st.free_stream_slots; }

static const uint8_t kAbc[32] = {0xba, 0x78, 0x16, 0xbf, 0x8f, 0x01, 0xcf, 0xea, 0x41, 0x41, 0x40, 0xde, 0x5d, 0xae, 0x22, 0x23,
                                 0xb0, 0x03, 0x61, 0xa3, 0x96, 0x17, 0x7a, 0x9c, 0xb4, 0x10, 0xff, 0x61, 0xf2, 0x00, 0x15, 0xad};
static const uint8_t kMillionA[32] = {0xcd, 0xc7, 0x6e, 0x5c, 0x99, 0x14, 0xfb, 0x92, 0x81, 0xa1, 0xc7, 0xe2, 0x84, 0xd7, 0x3e, 0x67,
                                      0xf1, 0x80, 0x9a, 0x48, 0xa4, 0x97, 0x20, 0x0e, 0x04, 0x6d, 0x39, 0xcc, 0xc7, 0x11, 0x2c, 0xd0};

int main(int argc, char **argv)
{
    const char *cas_dir = argc > 1 && argv[1][0] ? argv[1] : NULL;
    dm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.hbm_cas_bytes = 64u << 20; cfg.ring_bytes = 8u << 20; cfg.slab_bytes = 256u << 10; cfg.max_streams = 64;
    cfg.cas_dir = cas_dir;
    cfg.flags = DM_F_NUMA_LOCAL;
    if (dm_engine_create(&cfg, &E) != DM_OK) {
        char why[256]; size_t n = 0;
        dm_error_detail(NULL, 0, why, sizeof why, &n);
        fprintf(stderr, "dm_engine_create failed: %s\n", why);
        return 2;
    }
    call c;
    const size_t N = 1000000;
    uint8_t *a = (uint8_t *)malloc(N), *back = (uint8_t *)malloc(N);
    memset(a, 'a', N);

    /* 1. "abc": open / write / finish each on its own thread */
    memset(&c, 0, sizeof c); c.fn = c_open; c.u64 = 3; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_write; c.in = "abc"; c.len = 3; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_finish; on_new_thread(&c); CHECK(c.rc == DM_OK && c.matched == 1 && memcmp(c.dig, kAbc, 32) == 0);

    /* 2. one million 'a' with the expected digest, every 37 KiB write from a different thread, metadata, flush, finish */
    memset(&c, 0, sizeof c); c.fn = c_open; c.digest = kMillionA; c.u64 = N; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_meta; c.str = "\"etag-1\""; on_new_thread(&c); CHECK(c.rc == DM_OK);
    for (size_t off = 0; off < N; off += 37 * 1024) {
        c.fn = c_write; c.in = a + off; c.len = N - off < 37 * 1024 ? N - off : 37 * 1024; on_new_thread(&c); CHECK(c.rc == DM_OK);
    }
    c.fn = c_flush; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_finish; on_new_thread(&c); CHECK(c.rc == DM_OK && c.matched == 1 && memcmp(c.dig, kMillionA, 32) == 0);

    /* 3. served back, reads from different threads; sidecar; contains; alias by URL */
    memset(&c, 0, sizeof c); c.fn = c_contains; c.digest = kMillionA; on_new_thread(&c); CHECK(c.rc == DM_OK && c.u64 == N);
    c.fn = c_alias_put; c.str = "https://huggingface.co/x/resolve/main/a.bin"; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_alias_get; on_new_thread(&c); CHECK(c.rc == DM_OK && memcmp(c.dig, kMillionA, 32) == 0);
    c.fn = c_request; on_new_thread(&c); CHECK(c.rc == DM_OK && c.u64 == N);
    for (size_t off = 0; off < N;) {
        c.fn = c_cread; c.u64 = off; c.out = back + off; c.len = 100000; on_new_thread(&c); CHECK(c.rc == DM_OK && c.got > 0);
        if (c.rc != DM_OK || c.got == 0) break;
        off += c.got;
    }
    CHECK(memcmp(a, back, N) == 0);
    c.fn = c_cmeta; on_new_thread(&c); CHECK(c.rc == DM_OK && strstr(c.text, "etag-1") != NULL);
    c.fn = c_cclose; on_new_thread(&c); CHECK(c.rc == DM_OK);

    /* 4. an error on one thread, its text fetched on another */
    c.fn = c_cread; c.u64 = 0; c.len = 10; on_new_thread(&c); CHECK(c.rc == DM_EINVAL);      /* the reader was closed */
    { const uint64_t dead = c.id; memset(c.text, 0, sizeof c.text); c.fn = c_detail; c.id = dead; on_new_thread(&c);
      CHECK(c.rc == DM_OK && strstr(c.text, "unknown reader") != NULL); }

    /* 5. zero-copy window: acquire on one thread, fill, commit on another; range parts out of order; checkpoint / resume */
    memset(&c, 0, sizeof c); c.fn = c_open; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_acquire; on_new_thread(&c); CHECK(c.rc == DM_OK && c.len >= 3);
    if (c.rc == DM_OK) { memcpy(c.out, "abc", 3); c.len = 3; c.fn = c_commit; on_new_thread(&c); CHECK(c.rc == DM_OK); }
    c.fn = c_finish; on_new_thread(&c); CHECK(c.rc == DM_OK && memcmp(c.dig, kAbc, 32) == 0);
    c.fn = c_cevict; c.digest = kMillionA; on_new_thread(&c); CHECK(c.rc == DM_OK);
    memset(&c, 0, sizeof c); c.fn = c_open; c.digest = kMillionA; c.u64 = N; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_write_at; c.u64 = 600000; c.in = a + 600000; c.len = 400000; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_write_at; c.u64 = 0; c.in = a; c.len = 600000; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_finish; on_new_thread(&c); CHECK(c.rc == DM_OK && c.matched == 1);
    c.fn = c_cevict; on_new_thread(&c); CHECK(c.rc == DM_OK);
    memset(&c, 0, sizeof c); c.fn = c_open; c.digest = kMillionA; c.u64 = N; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_write; c.in = a; c.len = 500000; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_ckpt; on_new_thread(&c); CHECK(c.rc == DM_OK && c.ck.bytes == 500000 / 64 * 64);
    c.fn = c_abort; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_resume; c.u64 = N; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_write; c.in = a + c.ck.bytes; c.len = N - (size_t)c.ck.bytes; on_new_thread(&c); CHECK(c.rc == DM_OK);
    c.fn = c_finish; on_new_thread(&c); CHECK(c.rc == DM_OK && c.matched == 1 && memcmp(c.dig, kMillionA, 32) == 0);

    /* 6. follower attached from a foreign thread while the body arrives from others */
    memset(&c, 0, sizeof c); c.fn = c_open; c.digest = kMillionA; c.u64 = N; on_new_thread(&c); CHECK(c.rc == DM_OK);
    { call f; memset(&f, 0, sizeof f); f.fn = c_follow; f.digest = kMillionA; on_new_thread(&f); CHECK(f.rc == DM_OK);
      c.fn = c_write; c.in = a; c.len = N; on_new_thread(&c); CHECK(c.rc == DM_OK);
      f.fn = c_cread; f.u64 = 0; f.out = back; f.len = 300000; on_new_thread(&f); CHECK(f.rc == DM_OK && f.got > 0 && memcmp(back, a, f.got) == 0);
      c.fn = c_finish; on_new_thread(&c); CHECK(c.rc == DM_OK && c.matched == 1);
      f.fn = c_cread; f.u64 = N - 10; f.len = 100; on_new_thread(&f); CHECK(f.rc == DM_OK && f.got == 10);
      f.fn = c_cclose; on_new_thread(&f); CHECK(f.rc == DM_OK); }

    /* 7. suspend / resume_saved (disk tier only) */
    if (cas_dir) {
        memset(&c, 0, sizeof c); c.fn = c_open; c.digest = kMillionA; c.u64 = N; on_new_thread(&c); CHECK(c.rc == DM_OK);
        c.fn = c_write; c.in = a; c.len = 700001; on_new_thread(&c); CHECK(c.rc == DM_OK);
        c.fn = c_suspend; on_new_thread(&c); CHECK(c.rc == DM_OK && c.u64 == 700001 / 64 * 64);
        c.fn = c_resume_saved; on_new_thread(&c); CHECK(c.rc == DM_OK && c.u64 == 700001 / 64 * 64);
        c.fn = c_write; c.in = a + c.u64; c.len = N - (size_t)c.u64; on_new_thread(&c); CHECK(c.rc == DM_OK);
        c.fn = c_finish; on_new_thread(&c); CHECK(c.rc == DM_OK && c.matched == 1 && memcmp(c.dig, kMillionA, 32) == 0);
    }

    /* 8. nothing left open, no state slot lost */
    memset(&c, 0, sizeof c); c.fn = c_stats; on_new_thread(&c); CHECK(c.rc == DM_OK && c.u64 == 0 && c.got == 64);
    dm_engine_destroy(E);
    free(a); free(back);
    if (failures) { printf("CGO SHAPE FAILED (%d)\n", failures); return 1; }
    printf("CGO SHAPE OK\n");
    return 0;
}
