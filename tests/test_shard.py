"""Multi-GPU host logic on CPU: digest-prefix partition + the world_size-2
gloo gather of verdicts (the only cross-rank traffic the path has)."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

from demodel_b200.shard import BlobRef, ShardRouter, owner_of, owner_of_url

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _manifest(n=500, seed=0):
    rng = np.random.default_rng(seed)
    return [BlobRef(hashlib.sha256(bytes([i % 256, i // 256])).digest(), int(rng.integers(1, 1 << 20)), i)
            for i in range(n)]


def test_partition_is_exact_cover():
    blobs = _manifest()
    for world in (1, 2, 4, 8, 3):
        parts = ShardRouter(world).partition(blobs)
        assert sorted(b.index for p in parts for b in p) == list(range(len(blobs)))
        for r in range(world):
            assert [b.index for b in ShardRouter(world, r).mine(blobs)] == [b.index for b in parts[r]]


def test_partition_is_roughly_balanced():
    blobs = _manifest(4000)
    for world in (2, 4, 8):
        counts = [len(p) for p in ShardRouter(world).partition(blobs)]
        assert min(counts) > 0.8 * len(blobs) / world
        assert max(counts) < 1.2 * len(blobs) / world
        assert sum(ShardRouter(world).load(blobs)) == sum(b.size for b in blobs)


def test_unknown_digest_routes_by_url_and_is_stable():
    u = "https://huggingface.co/meta-llama/Meta-Llama-3-8B/resolve/main/model-00001-of-00004.safetensors"
    assert owner_of_url(u, 8) == owner_of_url(u, 8)
    assert 0 <= owner_of_url(u, 8) < 8
    assert owner_of(bytes(32), 8) == 0 and owner_of(b"\xff" * 32, 8) == 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from tests import _oracle
    from demodel_b200.shard import ShardRouter
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = _oracle.load()
        # every rank derives the same manifest, hashes only its own shard
        bodies = [orc.blob(0xDE40DE1, i, 0, 1000 + 37 * i).tobytes() for i in range(40)]
        refs = [BlobRef(hashlib.sha256(b).digest(), len(b), i) for i, b in enumerate(bodies)]
        router = ShardRouter(world, rank)
        local = [(b.index, orc.sha256(bodies[b.index]), orc.sha256(bodies[b.index]) == b.digest) for b in router.mine(refs)]
        merged = router.gather_verdicts(local)
        q.put((rank, len(local), merged))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather_of_verdicts():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    (r0, n0, merged0), (r1, n1, merged1) = out
    assert n0 + n1 == 40 and n0 > 0 and n1 > 0
    assert merged1 is None                       # only rank 0 receives
    assert [m[0] for m in merged0] == list(range(40))
    assert all(m[2] for m in merged0)
