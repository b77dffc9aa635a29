"""Arena-backed KV cache manager -- drop-in for ``src/specdec/cache/kv_cache_manager.py``.

Same class name and methods as the reference's ``SafeKVCacheManager`` (kv_cache_manager.py:16-519):
``reset``, ``set_batch_size``, ``set_active_indices``, ``filter_cache_layer``, ``filter_kv_cache``,
``update_base_cache``, ``update_draft_cache``, ``get_base_past_kv``, ``get_draft_past_kv``,
``set_sequence_metadata``, ``get_sequence_lengths``, ``realign_kv_cache``, ``get_kv_shapes``.

What changes underneath (B200 layout): each cache is a preallocated ``KVArena``
``[layers, B, H, C, D]`` instead of a dict of tensors regrown by ``torch.cat`` (:194-199, :268-273).
* growth = ONE ``sd_kv_append`` launch for all layers (O(new rows) bytes; the reference copies the
  whole cache per layer per step);
* rollback (``realign_kv_cache``, :353-479: slice / clone / pad / stack per sequence per layer) = the
  per-sequence lengths change and the dropped rows are zeroed -- the returned ``[B,H,max_len,D]`` views
  equal the reference's re-stacked tensors bit for bit;
* ``get_*_past_kv`` returns prefix VIEWS of the arena (no copy); filtering to a strict subset of the
  batch is the one O(L) operation left (``index_select``, as in the reference, :72-112) and it compacts
  the arena so that later appends keep hitting the fast path.
Only the row movement of the hot path runs in this repo's kernel; subset filtering and the zeroing of
rolled-back rows use torch tensor ops (plumbing on rare paths).
"""

from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional, Tuple

import torch

from kernels import cuda_ops

from .kv_arena import KVArena

logger = logging.getLogger(__name__)


class _ArenaCache:
    """One model's cache: an arena plus the host-known dense length (the reference's tensors are dense
    [B,H,L,D], so every sequence of a cache shares one L)."""

    def __init__(self, arena: KVArena, batch: int) -> None:
        self.arena, self.batch, self.length = arena, batch, 0
        self._ticket = torch.zeros(1, dtype=torch.int32, device=arena.device)

    def views(self, length: Optional[int] = None):
        n = self.length if length is None else length
        return tuple((self.arena.k[l][: self.batch, :, :n], self.arena.v[l][: self.batch, :, :n])
                     for l in range(self.arena.n_layers))


class SafeKVCacheManager:
    def __init__(self, device: str, capacity: int = 4096, growth: float = 2.0) -> None:
        self.device = device
        self.capacity, self.growth = capacity, growth
        self.base_cache: Optional[_ArenaCache] = None
        self.draft_cache: Optional[_ArenaCache] = None
        self.current_batch_size: int = 0
        self.active_indices: Optional[torch.Tensor] = None
        self.base_sequence_lengths: Dict[int, int] = {}
        self.draft_sequence_lengths: Dict[int, int] = {}
        self.batch_to_global_map: Dict[int, int] = {}

    # ---- state (reference: :41-70) ---------------------------------------------------------------
    def reset(self) -> None:
        self.base_cache = None
        self.draft_cache = None
        self.current_batch_size = 0
        self.active_indices = None
        self.base_sequence_lengths = {}
        self.draft_sequence_lengths = {}
        self.batch_to_global_map = {}

    def set_batch_size(self, batch_size: int) -> None:
        if self.current_batch_size != batch_size:
            self.reset()
            self.current_batch_size = batch_size

    def set_active_indices(self, active_indices) -> None:
        if isinstance(active_indices, torch.Tensor):
            self.active_indices = active_indices
        else:
            self.active_indices = torch.tensor(active_indices, dtype=torch.long, device=torch.device(self.device))

    # ---- filtering (reference: :72-147) ----------------------------------------------------------
    def filter_cache_layer(self, cache_tensor, active_indices=None):
        if cache_tensor is None:
            return None
        idx = active_indices if active_indices is not None else self.active_indices
        if idx is None or idx.numel() == 0:
            return None
        B = cache_tensor.shape[0]
        if idx.numel() == B and bool(torch.all(idx == torch.arange(B, device=idx.device))):
            return cache_tensor
        if int(idx.max()) >= B:
            logger.warning("KV cache: max_index %d >= batch_size %d, returning None", int(idx.max()), B)
            return None
        return cache_tensor.index_select(0, idx.to(cache_tensor.device))

    def filter_kv_cache(self, cache, active_indices=None):
        if cache is None:
            return None
        idx = active_indices if active_indices is not None else self.active_indices
        if idx is None or idx.numel() == 0:
            return None
        layers = cache.views() if isinstance(cache, _ArenaCache) else tuple(tuple(cache[i]) for i in sorted(cache.keys()))
        out = []
        for k, v in layers:
            fk, fv = self.filter_cache_layer(k, idx), self.filter_cache_layer(v, idx)
            if fk is None or fv is None:
                return None
            out.append((fk, fv))
        return tuple(out)

    # ---- growth (reference: :149-285) ------------------------------------------------------------
    def _new_cache(self, new_kv) -> _ArenaCache:
        k0 = new_kv[0][0]
        B, H, S, D = k0.shape
        cap = max(self.capacity, int(S * self.growth) + 64)
        arena = KVArena(len(new_kv), B, H, D, cap, 1, k0.dtype, k0.device)
        return _ArenaCache(arena, B)

    def _append(self, cache: _ArenaCache, new_kv) -> _ArenaCache:
        """All layers in one sd_kv_append launch: the offered rows are read in place from ``new_kv``."""
        arena = cache.arena
        S = new_kv[0][0].shape[2]
        if cache.length + S > arena.capacity:  # geometric regrowth (amortised O(1) per row)
            bigger = KVArena(arena.n_layers, cache.batch, arena.n_kv_heads, arena.head_dim,
                             int((cache.length + S) * self.growth) + 64, 1, arena.dtype, arena.device)
            bigger.k[:, :, :, : cache.length].copy_(arena.k[:, : cache.batch, :, : cache.length])
            bigger.v[:, :, :, : cache.length].copy_(arena.v[:, : cache.batch, :, : cache.length])
            grown = _ArenaCache(bigger, cache.batch)
            grown.length = cache.length
            cache, arena = grown, bigger
        ks = [kv[0].detach() for kv in new_kv]
        vs = [kv[1].detach() for kv in new_kv]
        same = all(t.stride() == ks[0].stride() and t.shape == ks[0].shape and t.stride(3) == 1 for t in ks + vs)
        if not same:
            ks, vs = [t.contiguous() for t in ks], [t.contiguous() for t in vs]
        es = arena.k.element_size()
        table = torch.tensor([[arena.k[l].data_ptr(), arena.v[l].data_ptr(), ks[l].data_ptr(), vs[l].data_ptr()]
                              for l in range(arena.n_layers)], dtype=torch.int64).to(arena.device, non_blocking=True)
        lens = torch.full((cache.batch,), cache.length, dtype=torch.int32, device=arena.device)
        cuda_ops.kv_arena_append_launch(
            table, arena.n_layers, cache.batch, arena.n_kv_heads, S, arena.row_bytes, arena.capacity,
            (arena.k.stride(1) * es, arena.k.stride(2) * es),
            (ks[0].stride(0) * es, ks[0].stride(1) * es, ks[0].stride(2) * es),
            lens, None, None, 0, None, False, cache.length, arena.overflow, cache._ticket)
        cache._keepalive = (ks, vs, table, lens)  # until the launch has run
        cache.length += S
        return cache

    def _update(self, cache: Optional[_ArenaCache], new_kv, active_tensor=None, full_batch=None) -> _ArenaCache:
        batch_dim = new_kv[0][0].shape[0]
        if cache is not None and cache.batch != batch_dim:
            kept = None
            if active_tensor is not None and cache.batch == full_batch and batch_dim == active_tensor.numel():
                # keep only the active sequences (reference :202-217): compact into a fresh arena once
                kept = self.filter_kv_cache(cache, active_tensor)
            cache = None  # any other mismatch: reset with the new rows (reference :218-229)
            if kept is not None and kept[0][0].shape[2] > 0:
                cache = self._append(self._new_cache(kept), kept)
        if cache is None:
            cache = self._new_cache(new_kv)
        return self._append(cache, new_kv)

    def update_base_cache(self, new_kv, active_indices, tokens_appended: Optional[List[int]] = None) -> None:
        if new_kv is None or len(new_kv) == 0:
            return
        new_kv = tuple((kv[0], kv[1]) for kv in new_kv)
        dev = new_kv[0][0].device
        active_tensor = active_indices if isinstance(active_indices, torch.Tensor) else torch.tensor(
            list(active_indices), dtype=torch.long, device=dev)
        self.base_cache = self._update(self.base_cache, new_kv, active_tensor, self.current_batch_size)
        if tokens_appended is not None and len(tokens_appended) == len(active_indices):
            for batch_idx, added in zip(list(active_indices), tokens_appended):
                gid = self.batch_to_global_map.get(int(batch_idx))
                if gid is not None:
                    self.base_sequence_lengths[gid] = self.base_sequence_lengths.get(gid, 0) + added

    def update_draft_cache(self, new_kv) -> None:
        if new_kv is None or len(new_kv) == 0:
            return
        self.draft_cache = self._update(self.draft_cache, tuple((kv[0], kv[1]) for kv in new_kv))

    def get_base_past_kv(self, active_indices=None):
        return self.filter_kv_cache(self.base_cache, active_indices)

    def get_draft_past_kv(self, active_indices=None):
        return self.filter_kv_cache(self.draft_cache, active_indices)

    # ---- metadata (reference: :299-351) ----------------------------------------------------------
    def set_sequence_metadata(self, global_sequence_ids, sequence_lengths, batch_to_global_map=None) -> None:
        if len(global_sequence_ids) != len(sequence_lengths):
            raise ValueError(f"global_sequence_ids ({len(global_sequence_ids)}) and sequence_lengths "
                             f"({len(sequence_lengths)}) must have same length")
        for gid, n in zip(global_sequence_ids, sequence_lengths):
            self.base_sequence_lengths[gid] = n
            self.draft_sequence_lengths[gid] = n
        self.batch_to_global_map = (batch_to_global_map if batch_to_global_map is not None
                                    else {i: g for i, g in enumerate(global_sequence_ids)})

    def get_sequence_lengths(self, global_sequence_ids, cache_type: str = "base") -> List[int]:
        d = self.base_sequence_lengths if cache_type == "base" else self.draft_sequence_lengths
        return [d.get(g, 0) for g in global_sequence_ids]

    # ---- rollback (reference: :353-479) ----------------------------------------------------------
    def realign_kv_cache(self, global_sequence_ids, new_sequence_lengths, cache_type: str = "base") -> None:
        if len(global_sequence_ids) != len(new_sequence_lengths):
            raise ValueError(f"global_sequence_ids ({len(global_sequence_ids)}) and new_sequence_lengths "
                             f"({len(new_sequence_lengths)}) must have same length")
        cache = self.base_cache if cache_type == "base" else self.draft_cache
        lengths = self.base_sequence_lengths if cache_type == "base" else self.draft_sequence_lengths
        for gid, n in zip(global_sequence_ids, new_sequence_lengths):
            lengths[gid] = n
        if cache is None:
            return
        if cache.batch != len(global_sequence_ids):
            logger.warning("Batch size mismatch: cache has %d, but %d sequences provided. Skipping realignment.",
                           cache.batch, len(global_sequence_ids))
            return
        arena = cache.arena
        new_max = max(new_sequence_lengths) if len(new_sequence_lengths) else 0
        if new_max > arena.capacity:
            raise ValueError("realign beyond the arena capacity")
        # rows [min(new_len_b, cur), max(cur, new_max)) of sequence b become zero; rows below stay in place
        cur = cache.length
        hi = max(cur, new_max)
        lo = min(min(new_sequence_lengths), cur) if len(new_sequence_lengths) else 0
        if hi > lo:
            keep = torch.tensor([min(int(n), cur) for n in new_sequence_lengths], dtype=torch.int64, device=arena.device)
            drop = torch.arange(lo, hi, device=arena.device)[None, :] >= keep[:, None]  # [B, hi-lo]
            m = drop[None, :, None, :, None]
            arena.k[:, : cache.batch, :, lo:hi].masked_fill_(m, 0)
            arena.v[:, : cache.batch, :, lo:hi].masked_fill_(m, 0)
        cache.length = new_max
        self.current_batch_size = len(global_sequence_ids)

    def get_kv_shapes(self, cache_type: str = "base") -> Dict[str, Any]:
        cache = self.base_cache if cache_type == "base" else self.draft_cache
        lengths = self.base_sequence_lengths if cache_type == "base" else self.draft_sequence_lengths
        if cache is None:
            return {"cache_exists": False, "num_layers": 0, "sequence_lengths": dict(lengths)}
        shapes = {f"layer_{i}": {"key_shape": list(k.shape), "value_shape": list(v.shape)} for i, (k, v) in enumerate(cache.views())}
        return {"cache_exists": True, "num_layers": cache.arena.n_layers, "shapes": shapes,
                "sequence_lengths": dict(lengths), "batch_size": self.current_batch_size}
