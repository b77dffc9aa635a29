"""KV-cache containers -- drop-in for ``src/specdec/cache/kv_types.py`` (reference: kv_types.py:14-156).

``KVCache`` keeps the reference's fields and methods; the tensors it holds are normally prefix VIEWS of a
``KVArena`` (no copy).  ``from_hf_output`` additionally accepts the ``DynamicCache`` objects that
``transformers`` >= 4.36 / 5.x return (the reference indexes ``past_key_values[0][0]`` and breaks on them:
SURVEY.md section 8c)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch


def _as_layer_tuple(past_key_values):
    layers = getattr(past_key_values, "layers", None)
    if layers is not None:  # transformers DynamicCache: .layers[i].keys / .values
        return tuple((l.keys, l.values) for l in layers)
    return tuple((kv[0], kv[1]) for kv in past_key_values)


@dataclass
class KVCache:
    past_key_values: Tuple[Tuple[torch.Tensor, torch.Tensor], ...]
    seq_len: int
    dtype: torch.dtype
    device: torch.device

    @classmethod
    def from_hf_output(cls, past_key_values) -> "KVCache":
        """reference: kv_types.py:33-60."""
        if past_key_values is None or len(past_key_values) == 0:
            raise ValueError("Cannot create KVCache from empty past_key_values")
        kv = _as_layer_tuple(past_key_values)
        first_key = kv[0][0]
        return cls(past_key_values=kv, seq_len=first_key.shape[2], dtype=first_key.dtype, device=first_key.device)

    def slice_prefix(self, length: int) -> "KVCache":
        """reference: kv_types.py:62-86 (views, no copy)."""
        if length > self.seq_len:
            raise ValueError(f"Cannot slice length {length} from cache with seq_len {self.seq_len}")
        sliced = tuple((k[:, :, :length, :], v[:, :, :length, :]) for k, v in self.past_key_values)
        return KVCache(past_key_values=sliced, seq_len=length, dtype=self.dtype, device=self.device)

    def to(self, device: torch.device) -> "KVCache":
        if self.device == device:
            return self
        moved = tuple((k.to(device), v.to(device)) for k, v in self.past_key_values)
        return KVCache(past_key_values=moved, seq_len=self.seq_len, dtype=self.dtype, device=device)

    def get_num_layers(self) -> int:
        return len(self.past_key_values)

    def get_shapes(self):
        if len(self.past_key_values) == 0:
            return ((), ())
        k, v = self.past_key_values[0]
        return (tuple(k.shape), tuple(v.shape))


def validate_kv_compatibility(base_cache: KVCache, new_cache: KVCache) -> None:
    """reference: kv_types.py:122-156."""
    if base_cache.get_num_layers() != new_cache.get_num_layers():
        raise ValueError(f"Layer count mismatch: base={base_cache.get_num_layers()}, new={new_cache.get_num_layers()}")
    if base_cache.dtype != new_cache.dtype:
        raise ValueError(f"Dtype mismatch: base={base_cache.dtype}, new={new_cache.dtype}")
    b, n = base_cache.get_shapes(), new_cache.get_shapes()
    if b[0][0] != n[0][0] or b[0][1] != n[0][1] or b[0][3] != n[0][3]:
        raise ValueError(f"Shape mismatch (excluding seq_len): base={b}, new={n}")
