// Host-side data structures of the engine that hold no CUDA state, kept in a
// header so they can be unit-tested natively on a CPU-only box
// (tests/native/test_host_util.cc, run by tests/test_native_host.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include <iterator>
#include <map>
#include <memory>
#include <utility>

namespace dm {

// First-fit free list over the HBM arena (byte offsets), coalescing on free.
class Arena {
public:
    void reset(uint64_t bytes) { free_.clear(); if (bytes) free_[0] = bytes; cap_ = bytes; used_ = 0; }
    bool alloc(uint64_t len, uint64_t *off)
    {
        for (auto it = free_.begin(); it != free_.end(); ++it) {
            if (it->second >= len) {
                *off = it->first;
                const uint64_t rest = it->second - len, at = it->first + len;
                free_.erase(it);
                if (rest) free_[at] = rest;
                used_ += len;
                return true;
            }
        }
        return false;
    }
    void release(uint64_t off, uint64_t len)
    {
        if (!len) return;
        used_ -= len;
        auto nx = free_.lower_bound(off);
        if (nx != free_.begin()) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == off) { off = pv->first; len += pv->second; free_.erase(pv); }
        }
        if (nx != free_.end() && off + len == nx->first) { len += nx->second; free_.erase(nx); }
        free_[off] = len;
    }
    uint64_t used() const { return used_; }
    uint64_t capacity() const { return cap_; }
    size_t fragments() const { return free_.size(); }
    const std::map<uint64_t, uint64_t> &free_map() const { return free_; }
private:
    std::map<uint64_t, uint64_t> free_;   // offset -> length of each free run
    uint64_t cap_ = 0, used_ = 0;
};

// Insert [lo, hi) into a set of disjoint half-open intervals (start -> end),
// merging with anything it touches or overlaps.
inline void add_interval(std::map<uint64_t, uint64_t> &m, uint64_t lo, uint64_t hi)
{
    if (lo >= hi) return;
    auto it = m.lower_bound(lo);
    if (it != m.begin()) {
        auto pv = std::prev(it);
        if (pv->second >= lo) { lo = pv->first; hi = hi > pv->second ? hi : pv->second; it = m.erase(pv); }
    }
    while (it != m.end() && it->first <= hi) { hi = hi > it->second ? hi : it->second; it = m.erase(it); }
    m[lo] = hi;
}

// Open-addressing index (linear probing, backward-shift deletion, at most half full) with the slice of the
// std::unordered_map interface the engine uses.  One 64-byte-class slot per entry holds the hash, the key and the value,
// so a lookup is ONE cache line whose address is known from the hash alone: batch operations over 10^5 digests
// prefetch() a few entries ahead instead of taking a chain of dependent misses per entry (a node-based map cost
// ~150 ns per lookup there).  Iterators are slot pointers: any insertion or erasure invalidates them.
template <class K, class V, class H>
class FlatIndex {
public:
    struct Slot { uint64_t h; K first; V second; };           // h == 0: empty
    using iterator = Slot *;
    FlatIndex() = default;
    FlatIndex(const FlatIndex &) = delete;
    FlatIndex &operator=(const FlatIndex &) = delete;
    iterator end() const { return nullptr; }
    size_t size() const { return n_; }
    size_t capacity() const { return cap_; }
    void prefetch(const K &k) const { if (cap_) __builtin_prefetch(&slots_[tag(k) & mask_]); }
    iterator find(const K &k) const
    {
        if (!cap_) return nullptr;
        const uint64_t h = tag(k);
        for (size_t i = h & mask_;; i = (i + 1) & mask_) {
            Slot &s = slots_[i];
            if (s.h == 0) return nullptr;
            if (s.h == h && s.first == k) return &s;
        }
    }
    void reserve(size_t n) { if (n * 2 > cap_) rehash(n * 2); }
    std::pair<iterator, bool> emplace(const K &k, V v)
    {
        if ((n_ + 1) * 2 > cap_) rehash((n_ + 1) * 2);
        const uint64_t h = tag(k);
        for (size_t i = h & mask_;; i = (i + 1) & mask_) {
            Slot &s = slots_[i];
            if (s.h == 0) { s.h = h; s.first = k; s.second = std::move(v); ++n_; return {&s, true}; }
            if (s.h == h && s.first == k) return {&s, false};
        }
    }
    V &operator[](const K &k) { return emplace(k, V()).first->second; }
    void erase(iterator it)
    {
        size_t i = (size_t)(it - slots_.get());
        for (;;) {                                           // close the gap: pull back whatever probed past slot i
            size_t j = i;
            bool found = false;
            for (;;) {
                j = (j + 1) & mask_;
                if (slots_[j].h == 0) break;
                const size_t home = slots_[j].h & mask_;
                if (((j - home) & mask_) >= ((j - i) & mask_)) { found = true; break; }
            }
            if (!found) break;
            slots_[i] = std::move(slots_[j]);
            i = j;
        }
        slots_[i].h = 0;
        slots_[i].second = V();
        --n_;
    }
    size_t erase(const K &k)
    {
        iterator it = find(k);
        if (!it) return 0;
        erase(it);
        return 1;
    }
    template <class F> void for_each(F f) const { for (size_t i = 0; i < cap_; ++i) if (slots_[i].h) f(slots_[i]); }
private:
    static uint64_t tag(const K &k) { const uint64_t h = (uint64_t)H()(k); return h ? h : 0x9E3779B97F4A7C15ull; }
    void rehash(size_t want)
    {
        size_t cap = 1024;
        while (cap < want) cap *= 2;
        if (cap <= cap_) return;
        std::unique_ptr<Slot[]> old(new Slot[cap]());
        old.swap(slots_);
        const size_t old_cap = cap_;
        cap_ = cap; mask_ = cap - 1;
        for (size_t i = 0; i < old_cap; ++i) {
            if (!old[i].h) continue;
            for (size_t j = old[i].h & mask_;; j = (j + 1) & mask_)
                if (slots_[j].h == 0) { slots_[j] = std::move(old[i]); break; }
        }
    }
    std::unique_ptr<Slot[]> slots_;
    size_t cap_ = 0, mask_ = 0, n_ = 0;
};

}  // namespace dm
