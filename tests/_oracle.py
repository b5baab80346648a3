"""Loader for the CPU oracle (oracle/): test infrastructure only.

Builds oracle/build/*.so with gcc on first use (seconds).  Nothing under
demodel_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")


class Oracle:
    def __init__(self):
        need = [os.path.join(ODIR, "build", n) for n in ("libdm_oracle.so", "libdm_cpu_baseline.so")]
        srcs = [os.path.join(ODIR, n) for n in ("sha256_oracle.c", "cpu_baseline.c")]
        stale = any(not os.path.exists(n) or os.path.getmtime(n) < os.path.getmtime(s) for n, s in zip(need, srcs))
        if stale:
            subprocess.run(["make", "-C", ODIR], check=True, capture_output=True)
        self.lib = C.CDLL(need[0])
        self.base = C.CDLL(need[1])
        self.lib.dmo_sha256.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        self.lib.dmo_sha256_chunked.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        self.lib.dmo_sha256_many.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        self.lib.dmo_sha256_init.argtypes = [C.c_void_p]
        self.lib.dmo_sha256_update.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self.lib.dmo_sha256_final.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.dmo_ctx_size.restype = C.c_uint
        self.lib.dmo_blob_fill.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]
        self.lib.dmo_blob_word.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        self.lib.dmo_blob_word.restype = C.c_uint64
        self.base.dmb_openssl_sha256.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        self.base.dmb_hash_and_cache.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t,
                                                 C.c_int, C.c_void_p]
        self.base.dmb_hash_and_cache.restype = C.c_double
        self.base.dmb_hash_and_cache_files.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint32, C.c_size_t,
                                                       C.c_int, C.c_void_p]
        self.base.dmb_hash_and_cache_files.restype = C.c_double

    @staticmethod
    def _arr(data) -> np.ndarray:
        if isinstance(data, np.ndarray):
            return np.ascontiguousarray(data.view(np.uint8))
        return np.frombuffer(bytes(data), dtype=np.uint8)

    def sha256(self, data) -> bytes:
        a = self._arr(data)
        out = np.zeros(32, dtype=np.uint8)
        self.lib.dmo_sha256(a.ctypes.data, a.nbytes, out.ctypes.data)
        return out.tobytes()

    def sha256_chunked(self, data, chunk: int) -> bytes:
        a = self._arr(data)
        out = np.zeros(32, dtype=np.uint8)
        self.lib.dmo_sha256_chunked(a.ctypes.data, a.nbytes, chunk, out.ctypes.data)
        return out.tobytes()

    def sha256_splits(self, data, cuts) -> bytes:
        """init / update at arbitrary split points / final."""
        a = self._arr(data)
        ctx = (C.c_uint8 * self.lib.dmo_ctx_size())()
        self.lib.dmo_sha256_init(ctx)
        prev = 0
        for c in list(cuts) + [a.nbytes]:
            self.lib.dmo_sha256_update(ctx, a.ctypes.data + prev, c - prev)
            prev = c
        out = np.zeros(32, dtype=np.uint8)
        self.lib.dmo_sha256_final(ctx, out.ctypes.data)
        return out.tobytes()

    def sha256_many(self, data, offsets) -> list[bytes]:
        a = self._arr(data)
        off = np.ascontiguousarray(np.asarray(offsets, dtype=np.uint64))
        n = len(off) - 1
        out = np.zeros(32 * max(n, 1), dtype=np.uint8)
        self.lib.dmo_sha256_many(a.ctypes.data, off.ctypes.data, n, out.ctypes.data)
        return [out[32 * i:32 * i + 32].tobytes() for i in range(n)]

    def openssl_sha256(self, data) -> bytes:
        a = self._arr(data)
        out = np.zeros(32, dtype=np.uint8)
        assert self.base.dmb_openssl_sha256(a.ctypes.data, a.nbytes, out.ctypes.data) == 0
        return out.tobytes()

    def blob(self, seed: int, blob: int, byte_off: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint8)
        self.lib.dmo_blob_fill(seed, blob, byte_off, out.ctypes.data, n)
        return out

    def hash_and_cache(self, src: np.ndarray, offsets, chunk: int = 32768, threads: int = 1, cache=None):
        """The CPU arm: returns (seconds, digests)."""
        off = np.ascontiguousarray(np.asarray(offsets, dtype=np.uint64))
        n = len(off) - 1
        out = np.zeros(32 * max(n, 1), dtype=np.uint8)
        cptr = cache.ctypes.data if cache is not None else None
        secs = self.base.dmb_hash_and_cache(src.ctypes.data, cptr, off.ctypes.data, n, chunk, threads, out.ctypes.data)
        return secs, [out[32 * i:32 * i + 32].tobytes() for i in range(n)]


    def hash_and_cache_files(self, src: np.ndarray, offsets, directory: str, chunk: int = 32768, threads: int = 1):
        """The CPU arm with the cache on a filesystem (tmpfs): returns (seconds, digests)."""
        off = np.ascontiguousarray(np.asarray(offsets, dtype=np.uint64))
        n = len(off) - 1
        out = np.zeros(32 * max(n, 1), dtype=np.uint8)
        os.makedirs(directory, exist_ok=True)
        secs = self.base.dmb_hash_and_cache_files(src.ctypes.data, directory.encode(), off.ctypes.data, n, chunk, threads,
                                                  out.ctypes.data)
        return secs, [out[32 * i:32 * i + 32].tobytes() for i in range(n)]


_inst = None


def load() -> Oracle:
    global _inst
    if _inst is None:
        _inst = Oracle()
    return _inst
