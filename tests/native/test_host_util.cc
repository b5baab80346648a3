// Native randomized tests of the engine's host-only data structures (no CUDA).
// Built and run by tests/test_native_host.py with plain g++.
#include "../../demodel_b200/csrc/host_util.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <unordered_map>
#include <vector>

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

static void arena_model_test(uint64_t seed)
{
    const uint64_t cap = 1 << 16, gran = 256;
    dm::Arena a;
    a.reset(cap);
    std::vector<uint8_t> owner(cap / gran, 0);               // brute-force model: which granules are taken
    struct Ext { uint64_t off, len; };
    std::vector<Ext> live;
    std::mt19937_64 rng(seed);
    uint64_t used = 0;
    for (int step = 0; step < 20000; ++step) {
        if (live.empty() || rng() % 3) {
            const uint64_t len = (1 + rng() % 24) * gran;
            uint64_t off = ~0ull;
            const bool ok = a.alloc(len, &off);
            // first fit: the model's first free run that is long enough
            uint64_t want = ~0ull, run = 0;
            for (uint64_t g = 0; g < owner.size(); ++g) {
                run = owner[g] ? 0 : run + 1;
                if (run * gran >= len) { want = (g + 1 - run) * gran; break; }
            }
            CHECK(ok == (want != ~0ull));
            if (ok) {
                CHECK(off == want && off % gran == 0 && off + len <= cap);
                for (uint64_t g = off / gran; g < (off + len) / gran; ++g) { CHECK(!owner[g]); owner[g] = 1; }
                live.push_back({off, len});
                used += len;
            }
        } else {
            const size_t k = rng() % live.size();
            Ext x = live[k];
            live[k] = live.back(); live.pop_back();
            if (rng() % 4 == 0 && x.len > gran) {                 // trim: free only the tail (publish() does this)
                const uint64_t keep = gran * (1 + rng() % (x.len / gran - 1 + 1));
                if (keep < x.len) {
                    a.release(x.off + keep, x.len - keep);
                    for (uint64_t g = (x.off + keep) / gran; g < (x.off + x.len) / gran; ++g) owner[g] = 0;
                    used -= x.len - keep;
                    x.len = keep;
                }
                live.push_back(x);
            } else {
                a.release(x.off, x.len);
                for (uint64_t g = x.off / gran; g < (x.off + x.len) / gran; ++g) owner[g] = 0;
                used -= x.len;
            }
        }
        CHECK(a.used() == used);
        // free map == maximal free runs of the model (fully coalesced, disjoint, sorted)
        if (step % 97 == 0) {
            auto it = a.free_map().begin();
            uint64_t g = 0;
            while (g < owner.size()) {
                if (owner[g]) { ++g; continue; }
                uint64_t h = g;
                while (h < owner.size() && !owner[h]) ++h;
                CHECK(it != a.free_map().end() && it->first == g * gran && it->second == (h - g) * gran);
                ++it; g = h;
            }
            CHECK(it == a.free_map().end());
        }
    }
    for (const Ext &x : live) a.release(x.off, x.len);
    CHECK(a.used() == 0 && a.fragments() == 1 && a.free_map().begin()->second == cap);
}

static void interval_model_test(uint64_t seed)
{
    const uint64_t span = 4096;
    std::mt19937_64 rng(seed);
    for (int round = 0; round < 300; ++round) {
        std::map<uint64_t, uint64_t> m;
        std::vector<uint8_t> bits(span, 0);
        const int n = 1 + (int)(rng() % 40);
        for (int i = 0; i < n; ++i) {
            uint64_t lo = rng() % span, hi = lo + rng() % 200;
            if (hi > span) hi = span;
            dm::add_interval(m, lo, hi);
            for (uint64_t b = lo; b < hi; ++b) bits[b] = 1;
            // the map must equal the maximal runs of set bits, where touching runs are merged
            auto it = m.begin();
            uint64_t b = 0;
            while (b < span) {
                if (!bits[b]) { ++b; continue; }
                uint64_t e = b;
                while (e < span && bits[e]) ++e;
                CHECK(it != m.end() && it->first == b && it->second == e);
                ++it; b = e;
            }
            CHECK(it == m.end());
        }
    }
    std::map<uint64_t, uint64_t> m;                           // empty and degenerate inputs
    dm::add_interval(m, 5, 5);
    dm::add_interval(m, 9, 3);
    CHECK(m.empty());
    dm::add_interval(m, 0, 10);
    dm::add_interval(m, 10, 20);                              // touching intervals merge
    CHECK(m.size() == 1 && m[0] == 20);
}

// FlatIndex against std::unordered_map under random insert / find / erase, with a hash that throws most keys into a
// few probe runs (long clusters, wrap-around at the end of the table) and with a well-spread one; values are
// shared_ptrs so that a slot dropped or moved wrongly shows up as a wrong use count.
struct Key32 { uint8_t b[32]; bool operator==(const Key32 &o) const { return memcmp(b, o.b, 32) == 0; } };
struct SpreadHash { size_t operator()(const Key32 &k) const { uint64_t v; memcpy(&v, k.b, 8); return (size_t)v; } };
struct ClusterHash { size_t operator()(const Key32 &k) const { uint64_t v; memcpy(&v, k.b, 8); return (size_t)(v % 7 == 0 ? 0 : 1016 + v % 13); } };
struct StdHash { size_t operator()(const Key32 &k) const { uint64_t v; memcpy(&v, k.b + 8, 8); return (size_t)v; } };

template <class H>
static void flat_index_model_test(uint64_t seed, int steps, int keyspace)
{
    dm::FlatIndex<Key32, std::shared_ptr<int>, H> idx;
    std::unordered_map<Key32, std::shared_ptr<int>, StdHash> model;
    std::mt19937_64 rng(seed);
    auto key_of = [&](uint64_t k) { Key32 x; std::mt19937_64 g(k * 7919 + 1); for (int i = 0; i < 32; i += 8) { const uint64_t w = g(); memcpy(x.b + i, &w, 8); } return x; };
    CHECK(idx.find(key_of(1)) == idx.end() && idx.erase(key_of(1)) == 0);
    for (int step = 0; step < steps; ++step) {
        const Key32 k = key_of(rng() % keyspace);
        const int op = (int)(rng() % 8);
        auto mit = model.find(k);
        if (op < 3) {
            auto v = std::make_shared<int>(step);
            auto r = idx.emplace(k, v);
            CHECK(r.second == (mit == model.end()));
            if (r.second) model.emplace(k, v);
            CHECK(*r.first->second == *model.find(k)->second && r.first->first == k);
        } else if (op < 5) {
            auto it = idx.find(k);
            CHECK((it == idx.end()) == (mit == model.end()));
            if (it != idx.end()) CHECK(it->second == mit->second && it->second.use_count() == 2);
            idx.prefetch(k);
        } else if (op == 5) {
            auto &slot = idx[k];                               // default-constructs when absent, like the map
            if (mit == model.end()) { CHECK(!slot); slot = std::make_shared<int>(-step); model[k] = slot; }
            else CHECK(slot == mit->second);
        } else if (op == 6) {
            auto it = idx.find(k);
            if (it != idx.end()) { idx.erase(it); model.erase(k); }
        } else {
            CHECK(idx.erase(k) == model.erase(k));
        }
        CHECK(idx.size() == model.size());
        if (step % 997 == 0 || step == steps - 1) {            // whole-table agreement
            size_t seen = 0;
            idx.for_each([&](const typename decltype(idx)::Slot &s) { ++seen; auto m = model.find(s.first); CHECK(m != model.end() && m->second == s.second); });
            CHECK(seen == model.size());
            for (auto &kv : model) { auto it = idx.find(kv.first); CHECK(it != idx.end() && it->second == kv.second && kv.second.use_count() == 2); }
            CHECK(idx.capacity() == 0 || idx.size() * 2 <= idx.capacity());
        }
    }
    if (keyspace <= 4096) {
        idx.reserve(100000);                                   // growth keeps every entry
        for (auto &kv : model) { auto it = idx.find(kv.first); CHECK(it != idx.end() && it->second == kv.second); }
    }
}

int main()
{
    for (uint64_t s = 1; s <= 3; ++s) { flat_index_model_test<SpreadHash>(s, 60000, 3000); flat_index_model_test<ClusterHash>(s, 40000, 700); }
    flat_index_model_test<SpreadHash>(9, 200000, 40000);
    for (uint64_t s = 1; s <= 4; ++s) arena_model_test(s);
    for (uint64_t s = 1; s <= 4; ++s) interval_model_test(s);
    printf("host_util ok\n");
    return 0;
}
