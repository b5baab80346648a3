#!/usr/bin/env python3
"""Randomized soak of the engine: many threads mixing every way of getting bytes in and out, for
SECONDS seconds, checking every digest / byte against hashlib and the leak accounting at the end.

    SECONDS=60 THREADS=12 python tools/soak.py        (needs a GPU)
"""
import ctypes as C
import hashlib
import os
import random
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import demodel_b200  # noqa: E402
from demodel_b200 import DM_ENOMEM, DM_ESTATE, DmError  # noqa: E402

SECONDS = float(os.environ.get("SECONDS", 60))
THREADS = int(os.environ.get("THREADS", 12))
SEED = 0xDE40DE1

rng0 = np.random.default_rng(1)
SIZES = [0, 1, 63, 64, 4096, 70000, 1 << 20, (1 << 20) + 7, 3 << 20, (5 << 20) + 123, 9 << 20]
BODIES = [demodel_b200.synth_fill_host(SEED, 10000 + i, 0, s) for i, s in enumerate(SIZES * 3)]
for i in range(len(SIZES), len(BODIES)):          # make the repeats distinct blobs
    if BODIES[i].nbytes:
        BODIES[i][0] ^= (i & 0xFF) | 1
WANT = [hashlib.sha256(b.tobytes()).digest() for b in BODIES]

errors, counts = [], {}
LIVE = {}          # thread id -> stream ids opened by the iteration in progress
lock = threading.Lock()


def note(k):
    with lock:
        counts[k] = counts.get(k, 0) + 1


class guard:
    """Abort the stream if the body of the with-block raises (a full arena under load is legal)."""
    def __init__(self, eng, sid):
        self.eng, self.sid = eng, sid

    def __enter__(self):
        return self.sid

    def __exit__(self, et, ev, tb):
        if et is not None:
            try:
                self.eng.stream_abort(self.sid)
            except DmError:
                pass
        return False


def worker(eng, tid, deadline):
    rnd = random.Random(tid)
    LIVE[tid] = []

    def t_open(*a, **k):
        sid = eng.stream_open(*a, **k)
        LIVE[tid].append(sid)
        return sid

    def t_resume(*a, **k):
        sid = eng.stream_resume(*a, **k)
        LIVE[tid].append(sid)
        return sid
    try:
        while time.time() < deadline:
            i = rnd.randrange(len(BODIES))
            body, want = BODIES[i], WANT[i]
            op = rnd.choice(["seq", "seq", "range", "zero", "abort", "mismatch", "read", "read", "follow", "evict", "resume", "unknown"])
            try:
                if op == "seq":
                    d, ok = eng.ingest(body, expect=want, chunk=rnd.choice([4096, 32768, 100000, 1 << 20]))
                    assert ok and d == want
                elif op == "unknown":                          # no Content-Length, no expected digest
                    d, ok = eng.ingest(body, expect=None, chunk=65536, size_hint=0)
                    assert ok and d == want
                elif op == "range" and body.nbytes > 8192:
                    sid = t_open(want, body.nbytes)
                    cuts = sorted(rnd.sample(range(1, body.nbytes), 3))
                    parts = list(zip([0] + cuts, cuts + [body.nbytes]))
                    pieces = []
                    for lo, hi in parts:
                        o = lo
                        while o < hi:
                            n = min(rnd.randrange(1, 300000), hi - o)
                            pieces.append((lo, o, n)); o += n
                    # interleave parts but keep each part's pieces in order
                    queues = {lo: [p for p in pieces if p[0] == lo] for lo, _ in parts}
                    while queues:
                        lo = rnd.choice(list(queues))
                        _, o, n = queues[lo].pop(0)
                        eng.stream_write_at(sid, o, body[o:o + n])
                        if not queues[lo]:
                            del queues[lo]
                    d, ok = eng.stream_finish(sid)
                    assert ok and d == want
                elif op == "zero":
                    sid = t_open(want, body.nbytes)
                    pos = 0
                    while pos < body.nbytes:
                        ptr, cap = eng.stream_acquire(sid)
                        n = min(cap, body.nbytes - pos, rnd.randrange(1, 200000))
                        C.memmove(ptr, body.ctypes.data + pos, n)
                        eng.stream_commit(sid, n)
                        pos += n
                    d, ok = eng.stream_finish(sid)
                    assert ok and d == want
                elif op == "abort":
                    sid = t_open(want, body.nbytes)
                    eng.stream_write(sid, body[:body.nbytes // 2])
                    eng.stream_abort(sid)
                elif op == "mismatch":
                    d, ok = eng.ingest(body, expect=bytes(32))
                    assert not ok and d == want
                elif op == "read":
                    opened = eng.cache_open(want)
                    if opened:
                        rid, size = opened
                        assert size == body.nbytes
                        if size:
                            off = rnd.randrange(size)
                            n = rnd.randrange(1, 400000)
                            assert eng.cache_read(rid, off, n) == body[off:off + n].tobytes()
                        eng.cache_close(rid)
                elif op == "follow":
                    f = eng.cache_follow(want)
                    if f:
                        rid, _ = f
                        got = bytearray()
                        try:
                            while True:
                                p = eng.cache_read(rid, len(got), 150000)
                                if not p:
                                    break
                                got += p
                            assert bytes(got) == body.tobytes()
                            note("follow_ok")
                        except DmError as ex:
                            assert ex.code in (DM_ESTATE, demodel_b200.DM_ENOENT), ex
                            assert bytes(got) == body.tobytes()[:len(got)]
                        eng.cache_close(rid)
                elif op == "evict":
                    try:
                        eng.cache_evict(want)
                    except DmError as ex:
                        assert ex.code == DM_ESTATE                # readers open
                elif op == "resume" and body.nbytes > 200000:
                    cut = rnd.randrange(100000, body.nbytes)
                    a = t_open(want, body.nbytes)
                    eng.stream_write(a, body[:cut])
                    ck, done = eng.stream_checkpoint(a)
                    assert done == cut // 64 * 64
                    eng.stream_abort(a)
                    b = t_resume(ck, expect=want, size_hint=body.nbytes)
                    if rnd.random() < 0.5:
                        eng.stream_write_at(b, 0, body[:done])
                    eng.stream_write(b, body[done:])
                    d, ok = eng.stream_finish(b)
                    assert ok and d == want
                note(op)
            except DmError as ex:
                if ex.code != DM_ENOMEM:                          # a full arena / ring under load is legal
                    raise
                for sid_ in list(LIVE.get(tid, [])):              # drop whatever this iteration left open
                    try:
                        eng.stream_abort(sid_)
                    except DmError:
                        pass
                note("enomem")
            finally:
                LIVE[tid] = []
    except Exception as ex:                                       # noqa: BLE001
        import traceback
        errors.append(traceback.format_exc())


def main():
    cas_dir = os.environ.get("CAS_DIR")                # e.g. /dev/shm/dm_soak: adds spill threads, disk-tier hits
    if cas_dir:
        import shutil
        shutil.rmtree(cas_dir, ignore_errors=True)
    eng = demodel_b200.Engine(device=0, hbm_cas_bytes=96 << 20, ring_bytes=24 << 20, slab_bytes=1 << 20, max_streams=4096,
                              cas_dir=cas_dir)
    deadline = time.time() + SECONDS
    ths = [threading.Thread(target=worker, args=(eng, t, deadline)) for t in range(THREADS)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(SECONDS + 120)
    alive = [t for t in ths if t.is_alive()]
    st = None
    for _ in range(300):
        st = eng.stats()
        if st["ring_slabs_free"] == st["ring_slabs_total"] and st["free_stream_slots"] == 4096:
            break
        time.sleep(0.01)
    print("ops:", dict(sorted(counts.items())))
    print("stats:", {k: st[k] for k in ("blobs_committed", "blobs_mismatched", "kernel_launches", "launches_deep", "ring_waits",
                                         "open_streams", "open_readers", "ring_slabs_free", "ring_slabs_total", "free_stream_slots",
                                         "hbm_cas_used")})
    ok = not errors and not alive and st["open_streams"] == 0 and st["open_readers"] == 0 \
        and st["ring_slabs_free"] == st["ring_slabs_total"] and st["free_stream_slots"] == 4096
    for e_ in errors[:3]:
        print(e_)
    if alive:
        print(f"{len(alive)} worker(s) hung")
    if cas_dir and ok:                                  # every file of the disk tier must hash to its own name
        bad = n_files = 0
        for root, _, files in os.walk(os.path.join(cas_dir, "blobs", "sha256")):
            for f in files:
                if f.endswith(".meta") or f.endswith(".part"):
                    continue
                n_files += 1
                if hashlib.sha256(open(os.path.join(root, f), "rb").read()).hexdigest() != f:
                    bad += 1
        print(f"disk tier: {n_files} files, {bad} with a wrong digest")
        ok = ok and bad == 0 and n_files > 0
    print("SOAK", "OK" if ok else "FAILED")
    if not alive:
        eng.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
