"""Device-side token validation -- drop-in for the reference's ``validate_and_clamp_tokens``
(src/specdec/utils/token_validation.py:15-85) without its host synchronisations.

The reference checks ``invalid_mask.any()`` (an implicit device->host sync) on every call and, when something is out
of range, reads min / max / count back for a log line: about twenty syncs per decoding step when its pipeline is the
caller.  Here the clamp is unconditional (it is the identity on valid ids, so the
returned values are the reference's in every case) and the evidence for the log line -- count, min and max of what was
clamped -- accumulates in a small device tensor per ``name``; ``validation_report()`` reads it back (ONE sync, when the
caller chooses) and logs what the reference would have logged.  ``strict=True`` keeps the reference's immediate report.
Same name, arguments and return value; ``get_vocab_size`` is unchanged."""
from __future__ import annotations

import logging
from typing import Dict, Optional

import torch

logger = logging.getLogger(__name__)

_I64_MAX, _I64_MIN = torch.iinfo(torch.int64).max, torch.iinfo(torch.int64).min
# name -> device int64 tensor [invalid_count, min_seen, max_seen, calls]
_stats: Dict[tuple, torch.Tensor] = {}


def _stat(name: str, device: torch.device) -> torch.Tensor:
    key = (name, str(device))
    t = _stats.get(key)
    if t is None:
        t = torch.tensor([0, _I64_MAX, _I64_MIN, 0], dtype=torch.int64, device=device)
        _stats[key] = t
    return t


def validate_and_clamp_tokens(input_ids: torch.Tensor, vocab_size: int, name: str = "input", strict: bool = False) -> torch.Tensor:
    """0 <= token < vocab_size for every returned id (reference: token_validation.py:15-76).  No host sync unless
    ``strict``; stays on the input's device; graph-capturable."""
    if input_ids is None or input_ids.numel() == 0:
        return input_ids
    ids64 = input_ids.to(torch.int64)
    invalid = (ids64 >= vocab_size) | (ids64 < 0)
    st = _stat(name, input_ids.device)
    st[0] += invalid.sum()
    st[1] = torch.minimum(st[1], ids64.min())
    st[2] = torch.maximum(st[2], ids64.max())
    st[3] += 1
    clamped = input_ids.clamp(min=0, max=vocab_size - 1)
    if strict:
        validation_report(name, vocab_size)
    return clamped


def validation_report(name: Optional[str] = None, vocab_size: Optional[int] = None, reset: bool = True) -> Dict[str, dict]:
    """Read the accumulated statistics back (one device->host copy per entry) and log out-of-range ids the way the
    reference does (token_validation.py:53-63).  Returns {name: {invalid, min, max, calls}}."""
    out = {}
    for (n, dev), t in list(_stats.items()):
        if name is not None and n != name:
            continue
        invalid, lo, hi, calls = (int(x) for x in t.tolist())
        out[n] = {"invalid": invalid, "min": lo if calls else None, "max": hi if calls else None, "calls": calls, "device": dev}
        if invalid:
            logger.error("[%s] Input ID out of bounds detected! Min: %s, Max: %s, Vocab_size: %s, Invalid_count: %s (over %s calls)",
                         n, lo, hi, vocab_size, invalid, calls)
        if reset:
            t.copy_(torch.tensor([0, _I64_MAX, _I64_MIN, 0], dtype=torch.int64))
    return out


def get_vocab_size(model) -> Optional[int]:
    """reference: token_validation.py:79-85."""
    if hasattr(model, "_model") and hasattr(model._model, "config"):
        return getattr(model._model.config, "vocab_size", None)
    return None
