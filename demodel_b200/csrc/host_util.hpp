// Host-side data structures of the engine that hold no CUDA state, kept in a
// header so they can be unit-tested natively on a CPU-only box
// (tests/native/test_host_util.cc, run by tests/test_native_host.py).
#pragma once
#include <cstdint>
#include <iterator>
#include <map>

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

}  // namespace dm
