"""dm_manifest_parse on the CPU (the reference's own manifest fixture) and
dm_manifest_prefetch on the GPU."""
import gzip
import json
import os

import numpy as np
import pytest

import demodel_b200
from demodel_b200 import manifest


def test_parse_reference_manifest(golden_dir):
    fx = json.load(open(os.path.join(golden_dir, "reference_fixture.json")))
    raw = gzip.decompress(bytes.fromhex(fx["gzip_body_hex"]))      # CONTRIBUTING.md:128-153
    layers = manifest.parse_manifest(raw)
    assert [(l.digest.hex(), l.size, l.media_type) for l in layers] == [
        (d["digest"].split(":")[1], d["size"], d["mediaType"]) for d in fx["manifest_layers"]]
    assert layers[1].size == 274290656 and layers[1].media_type == "application/vnd.ollama.image.model"


def test_parse_tolerates_nesting_and_rejects_garbage():
    d = "ab" * 32
    doc = {"schemaVersion": 2, "annotations": {"x": {"size": 3, "deep": [1, 2, {"k": None}]}},
           "config": {"digest": "sha256:" + d, "size": 7, "mediaType": "a/b", "annotations": {"size": "notanumber"}},
           "layers": [{"size": 1 << 40, "digest": "sha256:" + d.upper(), "platform": {"os": "linux"}},
                      {"digest": "md5:" + d, "size": 1}, {"digest": "sha256:" + d}, {"size": 1.5, "digest": "sha256:" + d}],
           "esc": 'q"uote\\ \u00e9 \n'}
    layers = manifest.parse_manifest(json.dumps(doc).encode())
    assert [(l.size, l.media_type) for l in layers] == [(7, "a/b"), (1 << 40, "")]
    assert layers[0].digest == layers[1].digest == bytes.fromhex(d)
    for bad in (b"", b"{", b'{"a":}', b'{"a":1,}', b'[1,2', b'{"a":1} trailing'):
        with pytest.raises(demodel_b200.DmError):
            manifest.parse_manifest(bad)
    assert manifest.parse_manifest(b"[]") == [] and manifest.parse_manifest(b"{}") == []


@pytest.mark.gpu
def test_prefetch_opens_verified_streams(oracle):
    bodies = [oracle.blob(0xDE40DE1, 700 + i, 0, n).tobytes() for i, n in enumerate((420, 300000, 11357, 17))]
    digs = [oracle.sha256(b) for b in bodies]
    doc = {"schemaVersion": 2, "config": {"mediaType": "application/vnd.docker.container.image.v1+json",
                                          "digest": "sha256:" + digs[0].hex(), "size": len(bodies[0])},
           "layers": [{"mediaType": "application/vnd.ollama.image.model", "digest": "sha256:" + d.hex(), "size": len(b)}
                      for d, b in zip(digs[1:], bodies[1:])] + [
                     {"mediaType": "dup", "digest": "sha256:" + digs[1].hex(), "size": len(bodies[1])}]}
    layers = manifest.parse_manifest(json.dumps(doc).encode())
    assert len(layers) == 5
    with demodel_b200.Engine(device=0, hbm_cas_bytes=64 << 20, ring_bytes=16 << 20) as eng:
        ids = manifest.prefetch(eng, layers)
        assert all(ids[:4]) and ids[4] == 0                  # the duplicate layer shares a stream
        assert eng.stats()["open_streams"] == 4
        for sid, body in zip(ids[:4], bodies):
            eng.stream_write(sid, body)
        for sid, d in zip(ids[:4], digs):
            got, ok = eng.stream_finish(sid)
            assert ok and got == d                           # verified against the manifest's digest
        assert manifest.prefetch(eng, layers) == [0] * 5     # second pull: everything is a hit
        # a corrupted layer body is caught by the digest the manifest promised
        eng.cache_evict(digs[2])
        ids = manifest.prefetch(eng, layers)
        assert [bool(i) for i in ids] == [False, False, True, False, False]
        eng.stream_write(ids[2], bodies[2][:-1] + b"X")
        got, ok = eng.stream_finish(ids[2])
        assert not ok and eng.cache_contains(digs[2]) is None


@pytest.mark.gpu
def test_ollama_pull_at_the_fixture_sizes(oracle, golden_dir):
    """BASELINE configs[3] in miniature: the reference's manifest (CONTRIBUTING.md:128-153) with its
    real layer sizes (one 274 MB model layer + three small blobs), synthetic bytes, digests substituted,
    sharded over however many GPUs are visible by digest prefix, pulled through pre-opened streams."""
    import hashlib
    import threading
    fx = json.load(open(os.path.join(golden_dir, "reference_fixture.json")))
    sizes = [l["size"] for l in fx["manifest_layers"]]
    assert sizes == [420, 274290656, 11357, 17]
    bodies = [demodel_b200.synth_fill_host(0xDE40DE1, 900 + i, 0, n) for i, n in enumerate(sizes)]
    digs = [hashlib.sha256(b.tobytes()).digest() for b in bodies]
    doc = json.loads(gzip.decompress(bytes.fromhex(fx["gzip_body_hex"])))
    doc["config"]["digest"] = "sha256:" + digs[0].hex()
    for layer, d in zip(doc["layers"], digs[1:]):
        layer["digest"] = "sha256:" + d.hex()
    layers = manifest.parse_manifest(json.dumps(doc).encode())
    assert [l.size for l in layers] == sizes and [l.digest for l in layers] == digs
    ngpu = demodel_b200.load().dm_device_count()
    engines = [demodel_b200.Engine(device=g, hbm_cas_bytes=512 << 20, ring_bytes=64 << 20) for g in range(ngpu)]
    try:
        home = [demodel_b200.shard_of(l.digest, ngpu) for l in layers]
        ids = {}
        for g, eng in enumerate(engines):
            mine = [l for l, h in zip(layers, home) if h == g]
            for l, sid in zip(mine, manifest.prefetch(eng, mine)):
                ids[l.digest] = (eng, sid)
        results = {}

        def pull(i):                                   # one goroutine per layer
            eng, sid = ids[digs[i]]
            for off in range(0, sizes[i], 1 << 20):
                eng.stream_write(sid, bodies[i][off:off + (1 << 20)])
            results[i] = eng.stream_finish(sid)
        ths = [threading.Thread(target=pull, args=(i,)) for i in range(4)]
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=120)
        assert all(results[i] == (digs[i], True) for i in range(4))
        eng, _ = ids[digs[1]]
        rid, size = eng.cache_open(digs[1])            # the model layer is a hit now, on its home GPU
        assert size == sizes[1] and eng.cache_read(rid, size - 4096, 4096) == bodies[1][-4096:].tobytes()
        eng.cache_close(rid)
    finally:
        for e in engines:
            e.close()


def test_parser_survives_mutated_and_random_input(golden_dir):
    """dm_manifest_parse eats bytes from the network: whatever it is fed it must return DM_OK or
    DM_EINVAL — never crash, hang or read out of bounds."""
    import ctypes as C
    import random
    from demodel_b200._lib import DM_EINVAL, DM_OK, DmLayer
    lib = demodel_b200.load()
    fx = json.load(open(os.path.join(golden_dir, "reference_fixture.json")))
    good = gzip.decompress(bytes.fromhex(fx["gzip_body_hex"]))
    rnd = random.Random(7)
    arr = (DmLayer * 8)()
    n = C.c_uint32()
    seen_ok = seen_bad = 0
    for it in range(4000):
        kind = it % 4
        if kind == 0:                                   # random bytes
            blob = bytes(rnd.getrandbits(8) for _ in range(rnd.randrange(0, 200)))
        elif kind == 1:                                 # truncation
            blob = good[:rnd.randrange(0, len(good))]
        elif kind == 2:                                 # byte flips
            b = bytearray(good)
            for _ in range(rnd.randrange(1, 6)):
                b[rnd.randrange(len(b))] = rnd.getrandbits(8)
            blob = bytes(b)
        else:                                           # structural noise: deep nesting, huge numbers, odd escapes
            blob = (b"[" * rnd.randrange(0, 200) + rnd.choice([b'{"size":1e999,"digest":"sha256:zz"}', b'"\\u12', b"-",
                    b'{"a":{"digest":"sha256:' + b"ab" * 32 + b'","size":18446744073709551616}}', good]) +
                    b"]" * rnd.randrange(0, 200))
        rc = lib.dm_manifest_parse(blob, len(blob), arr, 8, C.byref(n))
        assert rc in (DM_OK, DM_EINVAL), (rc, blob[:60])
        if rc == DM_OK:
            seen_ok += 1
            assert n.value <= 1000
        else:
            seen_bad += 1
    assert seen_ok > 50 and seen_bad > 500
    # max_layers smaller than what is found: count is reported, nothing is written past the array
    rc = lib.dm_manifest_parse(good, len(good), arr, 2, C.byref(n))
    assert rc == DM_OK and n.value == 4
