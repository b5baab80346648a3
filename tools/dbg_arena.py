import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, demodel_b200
n, sz = int(os.environ.get("N", 64)), int(os.environ.get("SZ", 64 << 20))
span = n * sz
eng = demodel_b200.Engine(device=0, hbm_cas_bytes=span + (1 << 30), ring_bytes=64 << 20)
offs = [i * sz for i in range(n)]; sizes = [sz] * n
dev = torch.empty(span, dtype=torch.uint8, device="cuda:0")
eng.synth_fill_device_many(0xDE40DE1, 0, dev.data_ptr(), offs, sizes)
digs, _, _ = eng.ingest_device(dev.data_ptr(), offs, sizes, hash_only=True)
exp = b"".join(digs)
print("learned", len(set(digs)), "distinct digests; stats", {k: v for k, v in eng.stats().items() if "hbm" in k})
for it in range(3):
    try:
        d, m, ms = eng.ingest_device(dev.data_ptr(), np.asarray(offs, dtype=np.uint64), np.asarray(sizes, dtype=np.uint64), expect=exp, replace=True, raw=True)
        print(it, bool(m.all()), ms, {k: v for k, v in eng.stats().items() if "hbm" in k or "commit" in k})
    except Exception as ex:
        print(it, "FAILED", ex, eng.stats())
        break
