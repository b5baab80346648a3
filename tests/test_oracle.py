"""Pins the CPU oracle (oracle/sha256_oracle.c) before anything trusts it.

The reference ships no test or golden vector for this path (SURVEY.md §4,
§8c), so the oracle is pinned against: the FIPS 180-4 / NIST CAVP known
answers, two independent implementations present in the image (Python
hashlib, OpenSSL EVP_sha256), and the one byte-level fixture the reference
holds (CONTRIBUTING.md:76-99).
"""
import gzip
import hashlib
import json
import os
import random

import numpy as np
import pytest


def _msg(spec):
    if "ascii" in spec:
        return spec["ascii"].encode()
    if "hex" in spec:
        return bytes.fromhex(spec["hex"])
    return spec["repeat"].encode() * spec["count"]


def test_fips_known_answers(oracle, golden_dir):
    vectors = json.load(open(os.path.join(golden_dir, "fips_vectors.json")))
    assert len(vectors) >= 8
    for v in vectors:
        m = _msg(v["message"])
        want = bytes.fromhex(v["sha256"])
        assert oracle.sha256(m) == want, v["name"]
        assert oracle.openssl_sha256(m) == want, v["name"]
        assert hashlib.sha256(m).digest() == want, v["name"]


def test_every_length_across_padding_boundaries(oracle):
    # 0..257 covers the one-vs-two padding block switch at 55/56, 63/64, 119/120
    rng = random.Random(1)
    for n in range(0, 258):
        m = bytes(rng.getrandbits(8) for _ in range(n))
        assert oracle.sha256(m) == hashlib.sha256(m).digest(), n


def test_streaming_splits_do_not_change_the_digest(oracle):
    rng = random.Random(2)
    m = bytes(rng.getrandbits(8) for _ in range(20000))
    want = hashlib.sha256(m).digest()
    for _ in range(50):
        k = rng.randint(0, 12)
        cuts = sorted(rng.randint(0, len(m)) for _ in range(k))
        assert oracle.sha256_splits(m, cuts) == want
    for chunk in (1, 63, 64, 65, 4096, 32768):
        assert oracle.sha256_chunked(m, chunk) == want


def test_large_random_vs_hashlib_and_openssl(oracle):
    m = np.random.default_rng(3).integers(0, 256, size=(8 << 20) + 13, dtype=np.uint8)
    want = hashlib.sha256(m.tobytes()).digest()
    assert oracle.sha256(m) == want
    assert oracle.openssl_sha256(m) == want
    assert oracle.sha256_chunked(m, 32768) == want


def test_reference_fixture(oracle, golden_dir):
    fx = json.load(open(os.path.join(golden_dir, "reference_fixture.json")))
    body = bytes.fromhex(fx["gzip_body_hex"])
    assert len(body) == fx["gzip_body_len"] == 379
    assert body[:2] == b"\x1f\x8b"                      # CONTRIBUTING.md:116 "file header 1f8b"
    assert oracle.sha256(body).hex() == fx["gzip_body_sha256"]
    assert oracle.openssl_sha256(body).hex() == fx["gzip_body_sha256"]
    raw = gzip.decompress(body)
    assert len(raw) == fx["inflated_len"] == 708
    assert oracle.sha256(raw).hex() == fx["inflated_sha256"]
    manifest = json.loads(raw)
    assert manifest["layers"][0]["size"] == 274290656   # CONTRIBUTING.md:141


def test_many_packed_blobs(oracle):
    rng = np.random.default_rng(4)
    sizes = [0, 1, 55, 56, 63, 64, 65, 119, 120, 1000, 4096, 70001]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    data = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    got = oracle.sha256_many(data, off)
    for i, s in enumerate(sizes):
        assert got[i] == hashlib.sha256(data[int(off[i]):int(off[i + 1])].tobytes()).digest()


def test_cpu_baseline_arm_matches(oracle):
    # the OpenSSL hash-and-cache loop timed by bench.py must itself be right
    rng = np.random.default_rng(5)
    sizes = [100000, 32768, 1, 0, 65536 + 7]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    src = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    cache = np.zeros_like(src)
    secs, digs = oracle.hash_and_cache(src, off, chunk=32768, threads=3, cache=cache)
    assert secs >= 0
    assert np.array_equal(cache, src)
    assert digs == oracle.sha256_many(src, off)


def test_blob_generator_is_deterministic_and_offset_consistent(oracle):
    seed = 0xDE40DE1
    a = oracle.blob(seed, 7, 0, 1000)
    assert np.array_equal(a, oracle.blob(seed, 7, 0, 1000))
    assert np.array_equal(a[123:777], oracle.blob(seed, 7, 123, 654))
    assert not np.array_equal(a, oracle.blob(seed, 8, 0, 1000))
    assert not np.array_equal(a, oracle.blob(seed + 1, 7, 0, 1000))
    # bytes look uniform enough that digests are not trivial
    big = oracle.blob(seed, 1, 0, 1 << 16)
    counts = np.bincount(big, minlength=256)
    assert counts.min() > 150 and counts.max() < 370


def test_cpu_baseline_file_variant(oracle, tmp_path):
    # BASELINE.md's form of the CPU arm: cache = content-addressed files
    rng = np.random.default_rng(6)
    sizes = [70000, 0, 32768, 5]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    src = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    secs, digs = oracle.hash_and_cache_files(src, off, str(tmp_path), threads=2)
    assert secs >= 0 and digs == oracle.sha256_many(src, off)
    for i, d in enumerate(digs):
        assert (tmp_path / d.hex()).read_bytes() == src[int(off[i]):int(off[i + 1])].tobytes()
