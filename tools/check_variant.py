#!/usr/bin/env python3
"""Parity + timing of one kernel variant pair, in a process of its own.

  tools/check_variant.py WIDE,DEEP        e.g.  tools/check_variant.py 9,4

Sets DM_KERNEL_VARIANT (read once, at engine creation; tuning only) and checks the deep kernel and the
2/4/8/16-streams-per-warp kernels - whose serial phase follows the DEEP number - and the wide kernel against
hashlib on: every length 0..300, group/line boundaries, 1500 ragged blobs in one launch, and the fused
CAS copy.  Then prints the kernel time of 256 x 8 MiB through the deep kernel next to the default variant's,
which is the number that decides whether a candidate ordering becomes the default.
Exit code 0 = bit-exact everywhere.
"""
import hashlib
import os
import sys

import numpy as np


def pack(blobs, align=16):
    offs, pos = [], 0
    for b in blobs:
        offs.append(pos)
        pos += (len(b) + align - 1) // align * align
    buf = np.zeros(max(pos, align), dtype=np.uint8)
    for o, b in zip(offs, blobs):
        buf[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    return buf, offs, [len(b) for b in blobs]


def run(variant):
    os.environ["DM_KERNEL_VARIANT"] = variant
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import demodel_b200
    e = demodel_b200.Engine(device=0, hbm_cas_bytes=3 << 30, ring_bytes=64 << 20, slab_bytes=1 << 20)
    rng = np.random.default_rng(41)
    lens = list(range(0, 301)) + [2047, 2048, 2049, 2048 + 55, 2048 + 56, 4096 + 63, 65535, 65536, 65537, 100003]
    msgs = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in lens]
    want = [hashlib.sha256(m).digest() for m in msgs]
    buf, offs, sizes = pack(msgs)
    dev = torch.from_numpy(buf).cuda()
    bad = 0
    for kernel in ("deep", 2, 4, 8, 16, "wide"):
        got, _, _ = e.ingest_device(dev.data_ptr(), offs, sizes, hash_only=True, kernel=kernel)
        n_bad = sum(g != w for g, w in zip(got, want))
        print(f"variant {variant} kernel {kernel!s:>5}: {len(msgs)} messages, {n_bad} mismatches")
        bad += n_bad
    # ragged, one launch
    rs = (rng.integers(0, 600, size=1500) * 16).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(rs)]).astype(np.uint64)
    data = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    ddev = torch.from_numpy(data).cuda()
    want = [hashlib.sha256(data[int(off[i]):int(off[i + 1])].tobytes()).digest() for i in range(len(rs))]
    for kernel in ("deep", 8):
        got, _, _ = e.ingest_device(ddev.data_ptr(), off[:-1], rs, hash_only=True, kernel=kernel)
        n_bad = sum(g != w for g, w in zip(got, want))
        print(f"variant {variant} kernel {kernel!s:>5}: 1500 ragged blobs, {n_bad} mismatches")
        bad += n_bad
    # fused CAS copy: hash-and-cache, then read back
    idx = [i for i in range(len(rs)) if rs[i] > 0][:64]
    got, matched, _ = e.ingest_device(ddev.data_ptr(), [int(off[i]) for i in idx], [int(rs[i]) for i in idx],
                                      expect=b"".join(want[i] for i in idx), replace=True, kernel="deep")
    bad += sum(g != want[i] for g, i in zip(got, idx)) + (0 if all(matched) else 1)
    i = max(idx, key=lambda j: rs[j])
    rid, sz = e.cache_open(want[i])
    back = e.cache_read(rid, 0, sz)
    e.cache_close(rid)
    if back != data[int(off[i]):int(off[i + 1])].tobytes():
        print("fused copy read-back differs")
        bad += 1
    # timing: 256 streams x 8 MiB, the deep kernel's regime
    n, size = 256, int(os.environ.get("CHECK_VARIANT_BLOB_KIB", 8192)) << 10
    big = torch.empty(n * size, dtype=torch.uint8, device="cuda")
    o, l = [i * size for i in range(n)], [size] * n
    e.synth_fill_device_many(0xDE40DE1, 0, big.data_ptr(), o, l)
    best = max(1e-6, min(e.ingest_device(big.data_ptr(), o, l, hash_only=True, kernel="deep", raw=True)[2] for _ in range(3)))
    print(f"variant {variant}: deep kernel 256 x {size >> 10} KiB  {best:.2f} ms  = {n * size / best / 1e6:.2f} GB/s "
          f"({size / best / 1e3:.1f} MB/s per stream)")
    e.close()
    return bad


if __name__ == "__main__":
    v = sys.argv[1] if len(sys.argv) > 1 else "9,4"
    bad = run(v)
    print("VARIANT OK" if bad == 0 else "VARIANT MISMATCH")
    sys.exit(0 if bad == 0 else 1)
