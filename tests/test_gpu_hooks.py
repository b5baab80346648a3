"""Round-2 hook-side behaviour through the C-ABI on the GPU (the same tests run on the CPU box against the
fake-runtime build, tests/native/run_mirror_tests.py): URL-keyed hits (HuggingFace resolve/ URLs carry no
digest), gzip-encoded manifests (the reference's own cached body, CONTRIBUTING.md:76-99), downloads suspended
to the disk tier and resumed after an engine restart, error text by id from a foreign thread, and the
finish / abort race.  Digests are checked against the oracle; bytes served back must be identical."""
import gzip
import hashlib
import json
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0xDE40DE1


@pytest.fixture(scope="module")
def eng():
    import demodel_b200
    e = demodel_b200.Engine(device=0, hbm_cas_bytes=1 << 30, ring_bytes=64 << 20, slab_bytes=1 << 20)
    yield e
    st = e.stats()
    assert st["open_streams"] == 0 and st["open_readers"] == 0, st
    e.close()


def _read_all(eng, rid, size, piece=32768):
    got = bytearray()
    while len(got) < size:
        part = eng.cache_read(rid, len(got), piece)
        assert part
        got += part
    return bytes(got)


def test_hf_url_without_a_digest_is_a_hit_the_second_time(eng, oracle):
    """start.go:197-200 hands the request hook a URL.  A HuggingFace resolve/ URL names no digest: the first
    fetch is ingested with the digest unknown, the verified body is indexed under its URL, and the second
    request for that URL is answered from the CAS, byte for byte."""
    url = "https://huggingface.co/meta-llama/Meta-Llama-3-8B/resolve/main/model-00001-of-00004.safetensors"
    body = oracle.blob(SEED, 8100, 0, (3 << 20) + 4321)
    assert eng.proxy_request(url) is None                              # miss: go upstream
    dig, ok = eng.proxy_fetch(url, body)                               # expect=None: digest learned by hashing
    assert ok and dig == oracle.sha256(body) == hashlib.sha256(body.tobytes()).digest()
    assert eng.alias_get(url) == dig
    hit = eng.proxy_request(url)
    assert hit is not None and hit[1] == body.nbytes
    assert _read_all(eng, hit[0], hit[1]) == body.tobytes()
    assert eng.cache_meta(hit[0])["headers"]["url"] == url             # the sidecar remembers where it came from
    eng.cache_close(hit[0])
    # same URL, new content upstream (a branch moved): the alias follows the newest verified body
    body2 = oracle.blob(SEED, 8101, 0, 70001)
    dig2, ok = eng.proxy_fetch(url, body2)
    assert ok and dig2 == oracle.sha256(body2) and eng.alias_get(url) == dig2
    hit = eng.proxy_request(url)
    assert _read_all(eng, hit[0], hit[1]) == body2.tobytes()
    eng.cache_close(hit[0])
    # a body that fails verification is not indexed
    url3 = url + "?rev=bad"
    _, ok = eng.proxy_fetch(url3, body2, expect=bytes(32))
    assert not ok and eng.alias_get(url3) is None and eng.proxy_request(url3) is None
    for d in (dig, dig2):
        eng.cache_evict(d)
    assert eng.proxy_request(url) is None                              # alias known, blob gone: a miss, not an error


def test_oci_url_names_its_digest(eng, oracle):
    body = oracle.blob(SEED, 8110, 0, 300001)
    dig = oracle.sha256(body)
    url = "https://registry.ollama.ai/v2/library/llama3/blobs/sha256:" + dig.hex()
    assert eng.proxy_request(url) is None
    got, ok = eng.proxy_fetch(url, body, expect=dig, chunk=4097)
    assert ok and got == dig
    hit = eng.proxy_request(url.replace("sha256:", "sha256-"))         # some mirrors spell it with a dash
    assert hit is not None and _read_all(eng, hit[0], hit[1], 100000) == body.tobytes()
    eng.cache_close(hit[0])
    assert eng.proxy_request("https://registry.ollama.ai/v2/library/llama3/blobs/sha256:" + "0" * 63) is None   # 63 hex: not a digest
    eng.cache_evict(dig)


def test_gzip_encoded_manifest_from_the_reference_fixture(eng, golden_dir):
    """The one cached body the reference documents (CONTRIBUTING.md:76-99) is a gzip-encoded Ollama manifest.
    Fed to the manifest hook as it arrived on the wire: inflated by dm_gunzip, parsed, one verified stream
    opened per layer."""
    import demodel_b200
    from demodel_b200.engine import gunzip
    fx = json.load(open(os.path.join(golden_dir, "reference_fixture.json")))
    wire = bytes.fromhex(fx["gzip_body_hex"])
    plain = gunzip(wire)
    assert plain == gzip.decompress(wire) and len(plain) == fx["inflated_len"]
    assert hashlib.sha256(plain).hexdigest() == fx["inflated_sha256"]
    with pytest.raises(demodel_b200.DmError):
        gunzip(wire[:-9] + bytes(9))                                   # damaged trailer
    with pytest.raises(demodel_b200.DmError):
        gunzip(wire, cap=100)                                          # too small: DM_ENOMEM
    rows = eng.proxy_manifest(wire, "gzip", chunk=100)                 # the body passes through in 100-byte reads
    want = [(d["digest"].split(":")[1], d["size"], d["mediaType"]) for d in fx["manifest_layers"]]
    assert [(r[0].hex(), r[1], r[2]) for r in rows] == want
    assert all(r[3] for r in rows)                                     # nothing cached yet: a stream per layer
    assert eng.stats()["open_streams"] == len(rows)
    for r in rows:
        eng.stream_abort(r[3])
    for enc in (None, "identity"):                                     # the same manifest, not encoded
        rows = eng.proxy_manifest(plain, enc)
        for r in rows:
            if r[3]:
                eng.stream_abort(r[3])
        assert [(r[0].hex(), r[1]) for r in rows] == [(w[0], w[1]) for w in want]
    with pytest.raises(demodel_b200.DmError):
        eng.proxy_manifest(wire, "br")                                 # an encoding we do not decode: reported, body untouched
    with pytest.raises(demodel_b200.DmError):
        eng.proxy_manifest(wire[:200], "gzip")                         # truncated on the wire


def test_suspended_download_survives_an_engine_restart(oracle, tmp_path):
    """an interrupted download is saved (SHA-256 mid-state + the bytes so far) under
    <cas_dir>/partial; a NEW engine over the same directory picks it up, the client resumes with
    Range: bytes=N-, the finished blob verifies against the oid and is cached whole."""
    import demodel_b200
    body = oracle.blob(SEED, 8200, 0, (9 << 20) + 777)
    want = oracle.sha256(body)
    cut = (5 << 20) + 12345
    cas = str(tmp_path / "cas")
    with demodel_b200.Engine(device=0, hbm_cas_bytes=256 << 20, ring_bytes=16 << 20, cas_dir=cas) as e1:
        assert e1.stream_resume_saved(want, body.nbytes) is None       # nothing saved yet
        sid = e1.stream_open(want, body.nbytes)
        for off in range(0, cut, 32768):
            e1.stream_write(sid, body[off:min(cut, off + 32768)])
        saved = e1.stream_suspend(sid)
        assert saved == cut // 64 * 64
        assert e1.stats()["suspended"] == 1 and e1.stats()["open_streams"] == 0
        with pytest.raises(demodel_b200.DmError):
            e1.stream_write(sid, body[:10])                            # the id was released
        # an engine without a disk tier cannot suspend
    with demodel_b200.Engine(device=0, hbm_cas_bytes=64 << 20, ring_bytes=8 << 20) as e0:
        sid = e0.stream_open(want, body.nbytes)
        e0.stream_write(sid, body[:100000])
        with pytest.raises(demodel_b200.DmError):
            e0.stream_suspend(sid)
        e0.stream_abort(sid)
    assert sorted(os.listdir(os.path.join(cas, "partial"))) == [want.hex() + ".ckpt", want.hex() + ".part"]
    with demodel_b200.Engine(device=0, hbm_cas_bytes=256 << 20, ring_bytes=16 << 20, cas_dir=cas) as e2:   # "proxy restart"
        assert e2.stats()["suspended"] == 1
        got = e2.stream_resume_saved(want, 0)                          # size comes from the saved record
        assert got is not None and got[1] == saved
        sid = got[0]
        for off in range(saved, body.nbytes, 100000):                  # upstream answers Range: bytes=<saved>-
            e2.stream_write(sid, body[off:off + 100000])
        dig, ok = e2.stream_finish(sid)
        assert ok and dig == want
        assert e2.fetch(want) == body.tobytes()                        # cached WHOLE: the saved prefix came back from disk
        assert e2.stats()["suspended"] == 0 and os.listdir(os.path.join(cas, "partial")) == []
        assert e2.stream_resume_saved(want, 0) is None


def test_url_alias_survives_an_engine_restart(oracle, tmp_path):
    import demodel_b200
    from demodel_b200._lib import DM_F_DISK_SYNC
    cas = str(tmp_path / "cas")
    url = "https://huggingface.co/org/repo/resolve/main/tokenizer.json"
    body = oracle.blob(SEED, 8300, 0, 123457)
    with demodel_b200.Engine(device=0, hbm_cas_bytes=64 << 20, ring_bytes=8 << 20, cas_dir=cas, flags=DM_F_DISK_SYNC) as e1:
        dig, ok = e1.proxy_fetch(url, body)
        assert ok and e1.stats()["aliases"] == 1
    with demodel_b200.Engine(device=0, hbm_cas_bytes=64 << 20, ring_bytes=8 << 20, cas_dir=cas) as e2:
        assert e2.stats()["aliases"] == 1 and e2.alias_get(url) == dig == oracle.sha256(body)
        hit = e2.proxy_request(url)                                    # served from the disk tier
        assert hit is not None and _read_all(e2, hit[0], hit[1]) == body.tobytes()
        e2.cache_close(hit[0])


def test_error_text_is_found_by_id_from_another_thread(eng):
    """cgo: a goroutine may sit on another OS thread when it asks why a call failed.  dm_last_error() is
    thread-local; dm_error_detail(e, id) is not."""
    import demodel_b200
    sid = eng.stream_open(None, 0)
    ptr, cap = eng.stream_acquire(sid)
    with pytest.raises(demodel_b200.DmError):
        eng.stream_write(sid, b"x")                                    # a window is outstanding
    seen = []
    t = threading.Thread(target=lambda: seen.append(eng.error_detail(sid)))
    t.start()
    t.join()
    assert seen and "not open for write" in seen[0]
    assert eng.error_detail(987654321) == ""                           # nothing failed on that id
    eng.stream_commit(sid, 0)
    eng.stream_abort(sid)


def test_finish_racing_abort_or_a_second_finish_never_hangs(eng, oracle):
    """a finish blocked on the final hash while another thread closes the stream used to wait for
    ever, and two finishes released the state slot twice."""
    import demodel_b200
    body = oracle.blob(SEED, 8400, 0, (2 << 20) + 5)
    want = oracle.sha256(body)
    free0 = eng.stats()["free_stream_slots"]
    for rnd in range(12):
        sid = eng.stream_open(want, body.nbytes)
        eng.stream_write(sid, body)
        out = {}

        def other():
            try:
                out["other"] = eng.stream_abort(sid) if rnd % 2 == 0 else eng.stream_finish(sid)
            except demodel_b200.DmError as ex:
                out["other"] = ex
        t = threading.Thread(target=other)
        t.start()
        try:
            out["main"] = eng.stream_finish(sid)
        except demodel_b200.DmError as ex:
            out["main"] = ex
        t.join(timeout=60)
        assert not t.is_alive(), "a finish or abort is stuck"
        for r in out.values():                                         # a verdict, or a clean "closed / unknown stream"
            assert isinstance(r, demodel_b200.DmError) or r is None or (r[1] and r[0] == want)
        eng.cache_evict(want)
    import time
    for _ in range(200):
        if eng.stats()["free_stream_slots"] == free0:
            break
        time.sleep(0.01)
    assert eng.stats()["free_stream_slots"] == free0                   # no slot leaked, none released twice
