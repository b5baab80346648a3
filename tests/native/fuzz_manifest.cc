// ASan/UBSan fuzz of dm_manifest_parse (the only parser of untrusted bytes on the path), built
// natively with g++ by tests/test_native_host.py.  The engine entry points manifest.cc references
// for dm_manifest_prefetch are stubbed: only the parser is exercised here.
#include "../../demodel_b200/csrc/manifest.cc"

#include <cstdio>
#include <random>

extern "C" int dm_cache_contains(dm_engine *, const uint8_t *, uint64_t *) { return DM_ENOENT; }
extern "C" int dm_stream_open(dm_engine *, const uint8_t *, uint64_t, uint64_t *) { return DM_ENODEV; }
extern "C" int dm_stream_abort(dm_engine *, uint64_t) { return DM_OK; }

int main()
{
    const std::string good =
        "{\"schemaVersion\":2,\"mediaType\":\"application/vnd.docker.distribution.manifest.v2+json\","
        "\"config\":{\"mediaType\":\"application/vnd.docker.container.image.v1+json\",\"digest\":\"sha256:"
        "31df23ea7daa448f9ccdbbcecce6c14689c8552222b80defd3830707c0139d4f\",\"size\":420},\"layers\":[{\"mediaType\":"
        "\"application/vnd.ollama.image.model\",\"digest\":\"sha256:970aa74c0a90ef7482477cf803618e776e173c007bf957f635f1015bfcfef0e6\","
        "\"size\":274290656},{\"mediaType\":\"x\",\"digest\":\"sha256:c71d239df91726fc519c6eb72d318ec65820627232b2f796219e87dcf35d0ab4\","
        "\"size\":11357,\"annotations\":{\"a\":[1,2,{\"b\":null}],\"esc\":\"q\\\"\\\\\\u00e9\"}}]}";
    std::mt19937_64 rng(11);
    dm_layer out[4];
    uint32_t n = 0;
    long ok = 0, bad = 0;
    if (dm_manifest_parse(good.data(), good.size(), out, 4, &n) != DM_OK || n != 3) { fprintf(stderr, "good manifest rejected\n"); return 1; }
    for (int it = 0; it < 200000; ++it) {
        std::string s;
        switch (it % 4) {
        case 0: s.resize(rng() % 160); for (auto &c : s) c = (char)rng(); break;
        case 1: s = good.substr(0, rng() % good.size()); break;
        case 2: s = good; for (int k = 0, m = 1 + (int)(rng() % 6); k < m; ++k) s[rng() % s.size()] = (char)rng(); break;
        default: s = std::string(rng() % 300, '[') + good.substr(rng() % good.size()) + std::string(rng() % 300, '}'); break;
        }
        // exact-size heap copy so any over-read trips ASan
        char *p = new char[s.size() ? s.size() : 1];
        if (!s.empty()) memcpy(p, s.data(), s.size());
        const int rc = dm_manifest_parse(p, s.size(), out, 4, &n);
        delete[] p;
        if (rc == DM_OK) ++ok; else if (rc == DM_EINVAL) ++bad; else { fprintf(stderr, "unexpected rc %d\n", rc); return 1; }
    }
    printf("manifest fuzz ok: %ld accepted, %ld rejected\n", ok, bad);
    return 0;
}
