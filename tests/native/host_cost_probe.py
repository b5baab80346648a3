"""TEST INFRASTRUCTURE: the engine's HOST-side cost per tiny body, measured on the CPU box against the fake runtime with
its kernels switched off (FAKE_CUDA_NULL_KERNEL=1: digests are meaningless, nothing is verified).  What remains is
stream open / write / finish, the pump, packs, completion threads, the index - the part of small-body throughput
that is not CUDA.   usage: host_cost_probe.py <fake libdemodel_b200.so> [threads ...]"""
import os
import sys
import time

os.environ["FAKE_CUDA_NULL_KERNEL"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import demodel_b200._lib as _lib  # noqa: E402

_lib.LIB_PATH = sys.argv[1]
import numpy as np  # noqa: E402
import demodel_b200  # noqa: E402

n, size = 200000, 4096
host = np.zeros(n * size, dtype=np.uint8)
off = np.arange(n + 1, dtype=np.uint64) * np.uint64(size)
with demodel_b200.Engine(device=0, hbm_cas_bytes=2 << 30, ring_bytes=256 << 20, slab_bytes=1 << 20, max_streams=65536) as eng:
    for th in [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8]:
        t0 = time.perf_counter()
        digs, ok, secs = eng.proxy_drive(host, off, chunk=32768, concurrency=4096, nthreads=th)
        st = eng.stats()
        print(f"{th:2d} threads: {n / secs / 1e3:7.0f} k bodies/s  ({secs / n * 1e6 * th:5.2f} us of thread time per body)  packed {st['packed_bodies']} in {st['packs']} packs, "
              f"launches {st['kernel_launches']}, ring waits {st['ring_waits']}")
        import ctypes as C
        lib = demodel_b200.load()
        for d in set(digs):
            lib.dm_cache_evict(eng._h, C.c_char_p(d))
