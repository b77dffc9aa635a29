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


class PeerExchange:
    """Fused multi-GPU result exchange over peer memory (include/sd_verify.h, ``sd_exchange``).

    The verify kernel's sequence finalizer stores each packed record into EVERY rank's exchange buffer
    (NVLink stores on CUDA-IPC mapped peer pointers), every value in one 8-byte word tagged with the epoch.
    The records are confirmed and gathered by the first CTA of the append that follows (fused launch or
    ``KVArena.append(exchange=...)``), or by ``wait()`` -- a 1-CTA kernel -- for steps without an append; either
    polls the tagged words, copies the table to ``out`` and advances the device-side epoch.  No NCCL launch on
    the step (a NCCL all-gather costs ~20 us on a ~33 us step at N=8).
    Pass ``desc_ptr`` to ``cuda_ops.verify_launch(exchange=...)`` / ``VerifyStep(exchange=...)``.

    One ``wait()`` per verify launch, on every rank, in the same order (SPMD); both are graph-capturable.

    ``lag=0``: a step gathers its own records, which couples the ranks step by step (each step costs the
    slowest rank's verify plus the NVLink latency).  ``lag=1``: a step gathers the PREVIOUS step's records:
    every rank still has its own shard's results immediately (``VerifyStep.packed``), the global table
    arrives one step later, and the ranks stop running in lockstep (a rank may be one step ahead)."""

    def __init__(self, shard: BatchShard, k_draft: int, device, group=None, lag: int = 0, _peers=None, _own=None) -> None:
        import ctypes as C

        from kernels import _capi

        if lag not in (0, 1):
            raise ValueError("lag must be 0 (gather the step's own records) or 1 (the previous step's)")
        self.lag = lag
        self.shard, self.K, self.width = shard, k_draft, k_draft + 2
        self.device = torch.device(device)
        self.rows = max(1, shard.max_local_batch)
        self._lib = lib = _capi.lib()
        world, rank = shard.world_size, shard.rank
        if world > _capi.MAX_PEERS:
            raise ValueError(f"at most {_capi.MAX_PEERS} ranks")
        self.nbytes = int(lib.sd_exchange_bytes(world, self.rows, self.width))
        self._opened = []
        self._owned = None
        with torch.cuda.device(self.device):
            if _peers is not None:  # single-process emulation: every buffer lives in this process
                peers = list(_peers)
            else:
                ptr, handle = C.c_void_p(), C.create_string_buffer(64)
                lib.check(lib.sd_exchange_alloc(self.nbytes, C.byref(ptr), handle), "sd_exchange_alloc")
                self._owned = ptr.value
                peers = [None] * world
                peers[rank] = ptr.value
                if world > 1:
                    handles = [None] * world
                    dist.all_gather_object(handles, handle.raw, group=group)
                    for r in range(world):
                        if r == rank:
                            continue
                        pp = C.c_void_p()
                        lib.check(lib.sd_exchange_open(handles[r], C.byref(pp)), f"sd_exchange_open(rank {r})")
                        peers[r] = pp.value
                        self._opened.append(pp.value)
            desc = _capi.ExchangeDesc()
            desc.world, desc.rank, desc.rows_per_rank, desc.width = world, rank, self.rows, self.width
            desc.total_rows = shard.global_batch
            desc.lag = lag
            for r in range(world):
                desc.peer[r] = peers[r]
            raw = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8)
            self.desc = raw.to(self.device)
            self.out = torch.empty((world * self.rows, self.width), dtype=torch.int32, device=self.device)
        self.peers = peers

    @property
    def desc_ptr(self) -> int:
        return self.desc.data_ptr()

    @classmethod
    def emulate(cls, world: int, global_batch: int, k_draft: int, device, lag: int = 0):
        """All ranks inside ONE process on one GPU (tests): the buffers are plain allocations, the peer
        tables point at each other.  Launch every rank's verify BEFORE any rank's wait()."""
        import ctypes as C

        from kernels import _capi

        lib = _capi.lib()
        shards = [BatchShard(global_batch, world, r) for r in range(world)]
        nbytes = int(lib.sd_exchange_bytes(world, max(1, shards[0].max_local_batch), k_draft + 2))
        ptrs = []
        with torch.cuda.device(device):
            for _ in range(world):
                ptr = C.c_void_p()
                lib.check(lib.sd_exchange_alloc(nbytes, C.byref(ptr), None), "sd_exchange_alloc")
                ptrs.append(ptr.value)
        out = [cls(s, k_draft, device, lag=lag, _peers=ptrs) for s in shards]
        out[0]._emulated_owned = ptrs  # freed with rank 0's object
        return out

    def wait(self) -> torch.Tensor:
        """Enqueue the wait kernel on the current stream; returns the gathered [global_batch, K+2] records
        (a view of a static buffer, valid once the stream reaches this point)."""
        with torch.cuda.device(self.device):
            rc = self._lib.sd_exchange_wait(self.desc_ptr, self.out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        self._lib.check(rc, "sd_exchange_wait")
        return self.unpad(self.out)

    def unpad(self, gathered: torch.Tensor) -> torch.Tensor:
        if self.shard.global_batch == self.rows * self.shard.world_size:
            return gathered
        parts = []
        for r in range(self.shard.world_size):
            s, e = self.shard.bounds(r)
            parts.append(gathered[r * self.rows : r * self.rows + (e - s)])
        return torch.cat(parts, dim=0)

    def close(self) -> None:
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for pp in self._opened:
                self._lib.sd_exchange_close(pp)
            self._opened = []
            if self._owned:
                self._lib.sd_exchange_free(self._owned)
                self._owned = None
            for pp in getattr(self, "_emulated_owned", []):
                self._lib.sd_exchange_free(pp)
            self._emulated_owned = []


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
