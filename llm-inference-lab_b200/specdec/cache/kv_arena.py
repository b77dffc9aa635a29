"""Preallocated KV-cache arena for every layer of one model, with device-side lengths.

B200 data layout for the reference's KV state (``KVCache`` tuples grown by ``torch.cat``:
src/specdec/cache/kv_types.py:14-31, src/specdec/cache/kv_cache_manager.py:194-199,:268-273):

    k, v      [n_layers, B, Hkv, C, D]   one allocation each; C = capacity in rows
    new_k/v   [n_layers, B, Hkv, R, D]   staging: the R = K+1 rows the verify forward produced
    seq_lens  [B] int32 (device)         rows valid per sequence -- rollback is a length update

One ``sd_kv_append`` launch (include/sd_verify.h) moves the accepted rows of every layer and
zeroes the rejected slots, reading ``accept_len`` straight from the verify kernel's output buffer:
O(R) bytes per (layer, sequence, head) instead of O(L), no host synchronisation
(the reference synchronises the device after every append: src/specdec/models/hf_wrappers.py:1021-1022).
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from kernels import cuda_ops


class KVArena:
    def __init__(
        self,
        n_layers: int,
        batch: int,
        n_kv_heads: int,
        head_dim: int,
        capacity: int,
        max_new_rows: int,
        dtype: torch.dtype = torch.bfloat16,
        device: str | torch.device = "cuda",
    ) -> None:
        self.n_layers, self.batch, self.n_kv_heads, self.head_dim = n_layers, batch, n_kv_heads, head_dim
        self.capacity, self.max_new_rows = capacity, max_new_rows
        self.dtype = dtype
        self.device = torch.device(device)
        shape = (n_layers, batch, n_kv_heads, capacity, head_dim)
        self.k = torch.zeros(shape, dtype=dtype, device=self.device)
        self.v = torch.zeros(shape, dtype=dtype, device=self.device)
        stage = (n_layers, batch, n_kv_heads, max_new_rows, head_dim)
        self.new_k = torch.zeros(stage, dtype=dtype, device=self.device)
        self.new_v = torch.zeros(stage, dtype=dtype, device=self.device)
        self.seq_lens = torch.zeros(batch, dtype=torch.int32, device=self.device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.layer_table = self._build_table(self.new_k, self.new_v)
        self._max_len_host = 0  # host-side upper bound of seq_lens, for the capacity check

    def step_state(self) -> torch.Tensor:
        """Device state of the fused step (include/sd_verify.h, sd_step_args.step_state): one 8-byte word per
        sequence + the launch epoch.  Zero once; the kernels never need it re-armed."""
        if getattr(self, "_step_state", None) is None:
            n = int(cuda_ops._capi.lib().sd_step_state_bytes(self.batch))
            self._step_state = torch.zeros((n + 7) // 8, dtype=torch.int64, device=self.device)
        return self._step_state

    # -- layout helpers -------------------------------------------------------------------------
    def use_staging(self, new_k: torch.Tensor, new_v: torch.Tensor, table: Optional[torch.Tensor] = None) -> None:
        """Point the arena at another pair of staged-row tensors [layers, B, Hkv, R, D] (e.g. the output buffers of
        the target forward); ``table`` = a prebuilt ``_build_table(new_k, new_v)`` (built here when omitted)."""
        assert new_k.shape == self.new_k.shape and new_v.shape == self.new_v.shape and new_k.stride() == new_v.stride()
        self.new_k, self.new_v = new_k, new_v
        self.layer_table = table if table is not None else self._build_table(new_k, new_v)

    def uniform_layers(self):
        """((arena_k0, arena_v0, new_k0, new_v0), arena layer stride, staging layer stride) in bytes when every
        layer sits at a fixed stride from layer 0 (this class always allocates that way), else None."""
        es = self.k.element_size()
        sa, sn = self.k.stride(0) * es, self.new_k.stride(0) * es
        if self.n_layers > 1 and (sa % 16 or sn % 16 or self.v.stride(0) * es != sa or self.new_v.stride(0) * es != sn):
            return None
        return (self.k.data_ptr(), self.v.data_ptr(), self.new_k.data_ptr(), self.new_v.data_ptr()), sa, sn

    def _build_table(self, new_k: torch.Tensor, new_v: torch.Tensor) -> torch.Tensor:
        """Device table of sd_kv_layer records.  The kernels issue 16-byte accesses on these pointers (the C ABI cannot
        look inside a device table), so alignment is checked HERE, where the pointers are known."""
        if self.row_bytes % 16 == 0:
            for name, t in (("arena k", self.k), ("arena v", self.v), ("staged k", new_k), ("staged v", new_v)):
                es = t.element_size()
                if t.data_ptr() % 16 or any((t.stride(d) * es) % 16 for d in range(t.dim() - 1)):
                    raise ValueError(f"{name}: base pointer and row/head/batch/layer strides must be 16-byte aligned "
                                     f"(ptr % 16 = {t.data_ptr() % 16}, strides {t.stride()}); pass a contiguous copy")
        rows = [
            [self.k[l].data_ptr(), self.v[l].data_ptr(), new_k[l].data_ptr(), new_v[l].data_ptr()]
            for l in range(self.n_layers)
        ]
        return torch.tensor(rows, dtype=torch.int64).to(self.device)

    @property
    def row_bytes(self) -> int:
        return self.head_dim * self.k.element_size()

    def layer(self, l: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Full-capacity [B,Hkv,C,D] views of layer ``l`` (what the attention kernel reads)."""
        return self.k[l], self.v[l]

    def past_key_values(self, length: int):
        """HF-style tuple of (k, v) [B,Hkv,length,D] views (reference: kv_types.py:62-86 slice_prefix)."""
        return tuple((self.k[l][:, :, :length], self.v[l][:, :, :length]) for l in range(self.n_layers))

    # -- state ------------------------------------------------------------------------------------
    def set_lengths(self, lengths, max_len_host: Optional[int] = None) -> None:
        lens = torch.as_tensor(lengths, dtype=torch.int32)
        if max_len_host is None:
            max_len_host = int(lens.max()) if lens.numel() else 0
        self.seq_lens.copy_(lens.to(self.device), non_blocking=True)
        self._max_len_host = max_len_host

    def reset(self) -> None:
        self.seq_lens.zero_()
        self.overflow.zero_()
        self._max_len_host = 0

    # -- the hot call -----------------------------------------------------------------------------
    def append(
        self,
        n_keep: Optional[torch.Tensor] = None,
        keep_bias: int = 0,
        zero_rejected: bool = True,
        accepted_mask: Optional[torch.Tensor] = None,
        rows: Optional[int] = None,
        check_capacity: bool = True,
        rows_ready_early: bool = False,
        exchange=None,
        exchange_packed: Optional[torch.Tensor] = None,
        seq_ready: Optional[torch.Tensor] = None,
    ) -> None:
        """Append the staged rows of every layer.  ``n_keep`` [B] int32 (device) + ``keep_bias`` rows
        of each sequence are kept (None = all ``rows``); lengths advance on device.
        ``rows_ready_early``: the staged rows were written at least two kernels ago on this stream
        (true in the verify step: target forward -> fused verify -> append), so the kernel may fetch
        them while the previous kernel drains (include/sd_verify.h, sd_kv_append).
        ``exchange``: a scheduler.batch_shard.PeerExchange whose records of this step are confirmed and
        gathered (into ``exchange.out``) by the first CTA of this launch -- replaces ``exchange.wait()``.
        ``exchange_packed``: this rank's packed [B, K+2] records; when given, the first CTA of this launch
        pushes them to the peers (then the verify launch must NOT carry the exchange).
        ``seq_ready``: the [B] int32 flag array that the verify launch just before received: the append then
        starts per sequence as soon as that sequence is verified (overlap with the verify tail)."""
        R = self.max_new_rows if rows is None else rows
        es = self.k.element_size()
        cuda_ops.kv_arena_append_launch(
            self.layer_table, self.n_layers, self.batch, self.n_kv_heads, R, self.row_bytes, self.capacity,
            (self.k.stride(1) * es, self.k.stride(2) * es),
            (self.new_k.stride(1) * es, self.new_k.stride(2) * es, self.new_k.stride(3) * es),
            self.seq_lens, self.seq_lens, n_keep, keep_bias, accepted_mask, zero_rejected,
            self._max_len_host if check_capacity else -1, self.overflow, self._ticket, rows_ready_early,
            exchange.desc_ptr if exchange is not None else None, exchange.out if exchange is not None else None,
            exchange_packed, seq_ready,
        )
        self._max_len_host = min(self.capacity, self._max_len_host + R)

    def step_bytes(self, mean_keep: float, rows: Optional[int] = None, zero_rejected: bool = True) -> int:
        """Algorithmic bytes of one append: accepted rows read, offered rows
        written (or accepted rows only when rejected slots are not zeroed)."""
        R = self.max_new_rows if rows is None else rows
        unit = self.n_layers * 2 * self.batch * self.n_kv_heads * self.row_bytes
        written = R if zero_rejected else mean_keep
        return int(unit * (mean_keep + written))
