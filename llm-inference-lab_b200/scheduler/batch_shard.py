"""Batch sharding of the verify step across the GPUs of one box.

The reference is single-device (no NCCL / torch.distributed call site anywhere; README lists
multi-GPU as planned) and its scheduler only juggles two streams
(src/scheduler/speculative_scheduler.py:79-92).  Sequences are independent in verify, kv_append and
attention (per-``b`` grids: src/kernels/cuda/verify.cu:41, kv_cache.cu:29-31), so the batch shards
with no data-path collective; the only exchange is the packed [B_local, K+2] int32 result record
(accept_len + K+1 target tokens) that every rank needs to advance its copy of the token sequences:
ONE all-gather per step (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class BatchShard:
    """Contiguous split of ``global_batch`` sequences over ``world_size`` ranks (rank r owns
    [start, stop)); sizes differ by at most one so any batch size shards."""

    global_batch: int
    world_size: int
    rank: int

    def __post_init__(self) -> None:
        if not (0 <= self.rank < self.world_size):
            raise ValueError("rank out of range")
        if self.global_batch < 0:
            raise ValueError("negative batch")

    def bounds(self, rank: Optional[int] = None) -> Tuple[int, int]:
        r = self.rank if rank is None else rank
        base, extra = divmod(self.global_batch, self.world_size)
        start = r * base + min(r, extra)
        return start, start + base + (1 if r < extra else 0)

    @property
    def start(self) -> int:
        return self.bounds()[0]

    @property
    def stop(self) -> int:
        return self.bounds()[1]

    @property
    def local_batch(self) -> int:
        s, e = self.bounds()
        return e - s

    @property
    def max_local_batch(self) -> int:
        return -(-self.global_batch // self.world_size)

    def take(self, t: torch.Tensor) -> torch.Tensor:
        """This rank's slice of a [global_batch, ...] tensor (a view)."""
        s, e = self.bounds()
        return t[s:e]


class StepResultExchange:
    """All-gather of the packed verify records.

    Buffers are preallocated so the collective can sit inside the captured CUDA graph of the step.
    Uneven shards are padded to ``max_local_batch`` rows (pad rows carry accept_len = -1)."""

    def __init__(self, shard: BatchShard, k_draft: int, device, group=None) -> None:
        self.shard, self.K, self.group = shard, k_draft, group
        self.width = k_draft + 2
        rows = shard.max_local_batch
        self.send = torch.full((rows, self.width), -1, dtype=torch.int32, device=device)
        self.recv = torch.empty((shard.world_size * rows, self.width), dtype=torch.int32, device=device)

    def all_gather(self, packed_local: torch.Tensor) -> torch.Tensor:
        """packed_local [B_local, K+2] int32 -> [global_batch, K+2] int32 on every rank."""
        n = self.shard.local_batch
        if packed_local.shape != (n, self.width):
            raise ValueError(f"expected packed [{n}, {self.width}], got {tuple(packed_local.shape)}")
        if self.shard.world_size == 1:
            return packed_local
        self.send[:n].copy_(packed_local, non_blocking=True)
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        return self.unpad(self.recv)

    def unpad(self, gathered: torch.Tensor) -> torch.Tensor:
        rows = self.shard.max_local_batch
        if self.shard.global_batch == rows * self.shard.world_size:
            return gathered
        parts = []
        for r in range(self.shard.world_size):
            s, e = self.shard.bounds(r)
            parts.append(gathered[r * rows : r * rows + (e - s)])
        return torch.cat(parts, dim=0)


def advance_sequences(tokens: torch.Tensor, lengths: torch.Tensor, packed: torch.Tensor, eos_id: int = -1):
    """Batched replacement of the per-sequence accept/bonus/append block of the reference
    (src/specdec/core/pipeline.py:3059 accepted = base_tokens[:accept_len]; :3140-3362 bonus token;
    :3572-3608 sequence growth) driven by the packed records: sequence b receives its accept_len[b]
    accepted tokens (== target ids on the accepted prefix) followed by the bonus token
    target_ids[b, accept_len[b]].

    tokens  [B, T] int64 token buffer, lengths [B] current lengths, packed [B, K+2] int32.
    Returns (new_lengths [B], n_new [B]).  Pure tensor ops on whatever device the inputs live on."""
    B, T = tokens.shape
    K = packed.shape[1] - 2
    acc = packed[:, 0].long()
    tgt = packed[:, 1:].long()
    n_new = acc + 1
    j = torch.arange(K + 1, device=tokens.device)[None, :]
    valid = j < n_new[:, None]
    if eos_id >= 0:  # truncate after the first EOS among the new tokens (pipeline.py:3120-3131)
        is_eos = (tgt == eos_id) & valid
        first = torch.where(is_eos.any(1), is_eos.float().argmax(1), torch.full_like(acc, K + 1))
        valid = valid & (j <= first[:, None])
        n_new = valid.sum(1)
    pos = lengths.long()[:, None] + j
    valid = valid & (pos < T)
    bidx = torch.arange(B, device=tokens.device)[:, None].expand_as(pos)
    tokens[bidx[valid], pos[valid]] = tgt[valid]
    n_new = valid.sum(1)
    return lengths + n_new.to(lengths.dtype), n_new
