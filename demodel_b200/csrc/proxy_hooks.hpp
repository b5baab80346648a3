// Host half of the path, above the C-ABI: what the reference's two goproxy
// hooks become once the engine is wired in.  The reference is Go and this
// image has no Go toolchain, so the wrappers are written in C++ with the same
// shape the Go ones have in go/demodel_b200.go (io.ReadCloser semantics), and
// the parity tests drive these instead.
//
//   BodyTee    — installed by the OnResponse hook around resp.Body
//                (/root/reference/cmd/demodel/start.go:201-204 returns resp
//                unchanged today).  Read(p) reads from upstream, tees the
//                bytes into the engine, returns them to goproxy's copy loop;
//                at EOF it finishes the stream and records the verdict;
//                Close() before EOF aborts (client went away / upstream error).
//   HitReader  — the body of the *http.Response the OnRequest hook returns to
//                short-circuit upstream (start.go:197-200 returns req,nil
//                today).  Read(p) serves bytes from the CAS.
#pragma once
#include "../../include/demodel_b200.h"

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace dm {

// io.Reader: returns bytes read, 0 at EOF, <0 on error.
struct Upstream {
    virtual ~Upstream() {}
    virtual long Read(void *p, size_t n) = 0;
};

class BodyTee {
public:
    BodyTee(dm_engine *e, Upstream *up, const uint8_t *expect, uint64_t content_length)
        : e_(e), up_(up)
    {
        rc_ = dm_stream_open(e, expect, content_length, &id_);
        open_ = rc_ == DM_OK;
    }
    ~BodyTee() { Close(); }

    // Reads up to n bytes from upstream into p (goproxy's buffer) and tees them.
    long Read(void *p, size_t n)
    {
        if (rc_ != DM_OK) return rc_;
        if (eof_) return 0;
        const long got = up_->Read(p, n);
        if (got < 0) { Abort(); return got; }
        if (got == 0) return Finish() == DM_OK ? 0 : rc_;
        rc_ = dm_stream_write(e_, id_, p, (size_t)got);
        if (rc_ != DM_OK) { Abort(); return rc_; }
        return got;
    }

    // Zero-copy form: the upstream read lands directly in the pinned ring;
    // *view points at the bytes for the client-side write.
    long ReadInPlace(const void **view, size_t max_n)
    {
        if (rc_ != DM_OK) return rc_;
        if (eof_) return 0;
        void *win = nullptr;
        size_t cap = 0;
        rc_ = dm_stream_acquire(e_, id_, &win, &cap);
        if (rc_ != DM_OK) { Abort(); return rc_; }
        const long got = up_->Read(win, cap < max_n ? cap : max_n);
        if (got < 0) { dm_stream_commit(e_, id_, 0); Abort(); return got; }
        rc_ = dm_stream_commit(e_, id_, (size_t)got);
        if (rc_ != DM_OK) { Abort(); return rc_; }
        if (got == 0) return Finish() == DM_OK ? 0 : rc_;
        *view = win;
        return got;
    }

    // Multiplexing drivers: Pump() moves one piece like Read()/ReadInPlace() but at EOF only
    // *starts* the final hash (dm_stream_flush); Wait() then collects the verdict.  A Go
    // goroutine would simply block in Read(); an OS thread playing many goroutines must not.
    long Pump(void *scratch, size_t n, bool in_place)
    {
        if (rc_ != DM_OK) return rc_;
        if (eof_ || flushed_) return 0;
        long got;
        if (in_place) {
            void *win = nullptr;
            size_t cap = 0;
            rc_ = dm_stream_acquire(e_, id_, &win, &cap);
            if (rc_ != DM_OK) { Abort(); return rc_; }
            got = up_->Read(win, cap < n ? cap : n);
            int rc2 = dm_stream_commit(e_, id_, got > 0 ? (size_t)got : 0);
            if (got >= 0 && rc2 != DM_OK) { rc_ = rc2; Abort(); return rc_; }
        } else {
            got = up_->Read(scratch, n);
            if (got > 0) {
                rc_ = dm_stream_write(e_, id_, scratch, (size_t)got);
                if (rc_ != DM_OK) { Abort(); return rc_; }
            }
        }
        if (got < 0) { Abort(); return got; }
        if (got == 0) {
            rc_ = dm_stream_flush(e_, id_);
            if (rc_ != DM_OK) { Abort(); return rc_; }
            flushed_ = true;
        }
        return got;
    }
    int Wait() { return eof_ ? rc_ : Finish(); }

    // io.Closer: before EOF this is an abort; after EOF a no-op.
    int Close()
    {
        if (open_ && !eof_) Abort();
        return DM_OK;
    }

    bool done() const { return eof_; }
    bool matched() const { return matched_ != 0; }
    const uint8_t *digest() const { return digest_; }
    int status() const { return rc_; }

private:
    int Finish()
    {
        eof_ = true;
        open_ = false;
        rc_ = dm_stream_finish(e_, id_, digest_, &matched_);
        if (rc_ != DM_OK) dm_stream_abort(e_, id_);     // some failures leave the stream open; a released id just says so (ids are never reused)
        return rc_;
    }
    void Abort()
    {
        if (open_) dm_stream_abort(e_, id_);
        open_ = false;
    }
    dm_engine *e_;
    Upstream *up_;
    uint64_t id_ = 0;
    int rc_ = DM_OK;
    bool open_ = false, eof_ = false, flushed_ = false;
    int matched_ = 0;
    uint8_t digest_[32] = {0};
};

class HitReader {
public:
    HitReader(dm_engine *e, const uint8_t digest[32]) : e_(e)
    {
        rc_ = dm_cache_open(e, digest, &id_, &size_);
        open_ = rc_ == DM_OK;
    }
    ~HitReader() { Close(); }
    bool hit() const { return open_; }
    uint64_t size() const { return size_; }     // Content-Length of the synthesised response
    long Read(void *p, size_t n)
    {
        if (!open_) return rc_;
        size_t got = 0;
        rc_ = dm_cache_read(e_, id_, off_, p, n, &got);
        if (rc_ != DM_OK) return rc_;
        off_ += got;
        return (long)got;                       // 0 == io.EOF
    }
    int Close()
    {
        if (open_) dm_cache_close(e_, id_);
        open_ = false;
        return DM_OK;
    }
private:
    dm_engine *e_;
    uint64_t id_ = 0, size_ = 0, off_ = 0;
    int rc_ = DM_OK;
    bool open_ = false;
};

}  // namespace dm
