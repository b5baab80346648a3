#!/usr/bin/env python3
"""Throughput of very small blobs through the stream API (manifests, configs, tokenizer files)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import demodel_b200
from tests import _oracle
n, size = int(os.environ.get("N", 100000)), int(os.environ.get("SIZE", 4096))
host = np.frombuffer(os.urandom(n * size), dtype=np.uint8)
off = np.arange(n + 1, dtype=np.uint64) * np.uint64(size)
with demodel_b200.Engine(device=0, hbm_cas_bytes=2 << 30, ring_bytes=1 << 30, slab_bytes=int(os.environ.get("SLAB", 1 << 20)),
                         max_streams=65536) as eng:
    for conc, threads in ((256, 8), (1024, 8), (4096, 8), (4096, 16)):
        digs, ok, secs = eng.proxy_drive(host, off, chunk=32768, concurrency=conc, nthreads=threads)
        assert all(ok)
        st = eng.stats()
        print(f"{n} x {size} B, {conc} concurrent on {threads} threads: {n / secs / 1e3:.0f} k blobs/s, {n * size / secs / 1e9:.2f} GB/s, "
              f"launches so far {st['kernel_launches']}, ring waits {st['ring_waits']}")
        for d in digs[:: max(1, n // 2000)]:
            pass
        lib = demodel_b200.load()
        import ctypes as C
        for d in digs:
            lib.dm_cache_evict(eng._h, C.c_char_p(d))
    want = _oracle.load().sha256(host[:size])
    assert digs[0] == want
