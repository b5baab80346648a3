#!/usr/bin/env python3
"""Disk-tier throughput: ingest N blobs resident in HBM with a cas_dir on tmpfs and wait until every
blob is on disk (spill threads: D2H on side streams + file writes), then serve them back from disk."""
import os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import demodel_b200

n, size = int(os.environ.get("N", 64)), int(os.environ.get("SIZE", 64 << 20))
root = "/dev/shm/dm_cas_probe"
shutil.rmtree(root, ignore_errors=True)
eng = demodel_b200.Engine(device=0, hbm_cas_bytes=n * size + (1 << 30), ring_bytes=256 << 20, cas_dir=root)
offs, lens = [i * size for i in range(n)], [size] * n
dev = torch.empty(n * size, dtype=torch.uint8, device="cuda")
eng.synth_fill_device_many(0xDE40DE1, 0, dev.data_ptr(), offs, lens)
t0 = time.perf_counter()
digs, ok, kms = eng.ingest_device(dev.data_ptr(), offs, lens)
t1 = time.perf_counter()
paths = [os.path.join(root, "blobs", "sha256", d.hex()[:2], d.hex()) for d in digs]
while not all(os.path.exists(p) for p in paths):
    time.sleep(0.005)
t2 = time.perf_counter()
print(f"hash+cache {n}x{size >> 20} MiB: {n * size / (t1 - t0) / 1e9:.1f} GB/s; spill to tmpfs complete after {t2 - t0:.2f} s "
      f"-> {n * size / (t2 - t0) / 1e9:.1f} GB/s sustained to disk tier")
for d in digs:
    eng.cache_evict(d)                      # HBM copies gone: the next reads come from the disk tier
out = np.empty(n * size, dtype=np.uint8)
hoff = np.arange(n + 1, dtype=np.uint64) * np.uint64(size)
secs = eng.proxy_serve(digs, out, hoff, chunk=1 << 20, nthreads=8)
import hashlib
assert hashlib.sha256(out[:size].tobytes()).digest() == digs[0]
print(f"serve from disk tier (pread, 8 threads): {n * size / secs / 1e9:.1f} GB/s")
eng.close()
shutil.rmtree(root, ignore_errors=True)
