"""Digest-prefix sharding across GPUs.

Blobs are independent and a single blob never spans GPUs (SHA-256 chains
block to block), so the multi-GPU path is a pure partition: every rank
computes the same owner for every digest and handles only its own.  There is
no data-path collective; the only cross-rank traffic is the optional gather of
per-rank verdict summaries (control plane, a few bytes per blob).

Blobs whose digest is not known up front (no LFS oid / OCI digest in the URL)
are routed by a hash of the URL instead and keep that home.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence


def owner_of(digest: bytes, n_shards: int) -> int:
    """Pure-Python twin of dm_shard_of (include/demodel_b200.h): the leading
    16 bits of the digest scaled onto [0, n_shards)."""
    if n_shards <= 1:
        return 0
    prefix = (digest[0] << 8) | digest[1]
    return (prefix * n_shards) >> 16


def owner_of_url(url: str, n_shards: int) -> int:
    """Home for a blob whose content digest is unknown before the body arrives."""
    return owner_of(hashlib.blake2s(url.encode(), digest_size=32).digest(), n_shards)


@dataclass(frozen=True)
class BlobRef:
    digest: bytes          # expected SHA-256 (32 bytes)
    size: int
    index: int = 0         # position in the caller's manifest


class ShardRouter:
    """Maps a manifest of blobs onto `world_size` engines, one per GPU."""

    def __init__(self, world_size: int, rank: int = 0):
        if world_size < 1 or not (0 <= rank < world_size):
            raise ValueError("bad world_size / rank")
        self.world_size = world_size
        self.rank = rank

    def owner(self, digest: bytes) -> int:
        return owner_of(digest, self.world_size)

    def mine(self, blobs: Iterable[BlobRef]) -> list[BlobRef]:
        return [b for b in blobs if self.owner(b.digest) == self.rank]

    def partition(self, blobs: Iterable[BlobRef]) -> list[list[BlobRef]]:
        parts: list[list[BlobRef]] = [[] for _ in range(self.world_size)]
        for b in blobs:
            parts[self.owner(b.digest)].append(b)
        return parts

    def load(self, blobs: Iterable[BlobRef]) -> list[int]:
        """Bytes per shard: the imbalance a digest-prefix split gives this manifest."""
        out = [0] * self.world_size
        for b in blobs:
            out[self.owner(b.digest)] += b.size
        return out

    def gather_verdicts(self, local: Sequence[tuple[int, bytes, bool]], group=None
                        ) -> Optional[list[tuple[int, bytes, bool]]]:
        """Collect (manifest index, digest, matched) from every rank on rank 0.

        Control plane only.  Uses torch.distributed if it is initialised
        (gloo on CPU, nccl on GPUs); with a single process it is the identity.
        """
        import torch.distributed as dist
        if self.world_size == 1 or not dist.is_available() or not dist.is_initialized():
            return sorted(local)
        gathered = [None] * self.world_size if dist.get_rank(group) == 0 else None
        dist.gather_object(list(local), gathered, dst=0, group=group)
        if gathered is None:
            return None
        merged = [v for part in gathered for v in part]
        return sorted(merged)
