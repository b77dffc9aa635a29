"""Kernel registry: op name -> candidate implementations ranked by priority, filtered by device.

Same public surface as the reference's ``src/kernels/registry.py:11-123`` (``register``,
``get_best``, ``list_available``, ``get_status``, the class-level ``_kernels`` table that the
reference's tests clear directly, and the module-level ``registry`` instance), so callers such as
``src/specdec/policies/policies.py:128`` and ``src/specdec/models/hf_wrappers.py:927-929`` work
unchanged.  Differences: registration is guarded by a lock (the scheduler drives two CUDA streams
from worker threads) and re-registering the same callable for the same op/device replaces the
entry instead of duplicating it.
"""

from __future__ import annotations

import logging
import threading
from typing import Any, Callable, Dict, List, Optional

logger = logging.getLogger(__name__)

_WILDCARD = "auto"


class KernelRegistry:
    """Priority/device dispatch table.  State is class-level, as in the reference (registry.py:14)."""

    _kernels: Dict[str, List[Dict[str, Any]]] = {}
    _lock = threading.Lock()

    @classmethod
    def register(cls, op_name: str, impl: Callable, priority: int, device: str = _WILDCARD) -> None:
        """Add ``impl`` for ``op_name``; higher ``priority`` wins; ``device`` in {"cuda","mps","cpu","auto"}."""
        entry = {"function": impl, "priority": int(priority), "device": device}
        with cls._lock:
            bucket = cls._kernels.setdefault(op_name, [])
            bucket[:] = [e for e in bucket if not (e["function"] is impl and e["device"] == device)]
            bucket.append(entry)
            # stable sort: among equal priorities the earlier registration stays first
            bucket.sort(key=lambda e: -e["priority"])
        logger.debug("registered %s -> %s (priority=%s, device=%s)", op_name, _name(impl), priority, device)

    @classmethod
    def _candidates(cls, op_name: str, device: str) -> List[Dict[str, Any]]:
        return [e for e in cls._kernels.get(op_name, ()) if e["device"] in (device, _WILDCARD)]

    @classmethod
    def get_best(cls, op_name: str, device: str) -> Optional[Callable]:
        """Highest-priority implementation usable on ``device`` (or registered for "auto"), else None."""
        if op_name not in cls._kernels:
            return None
        cands = cls._candidates(op_name, device)
        if not cands:
            logger.debug("no %s kernel registered for device %s", op_name, device)
            return None
        fn = cands[0]["function"]
        return fn if callable(fn) else None

    @classmethod
    def list_available(cls, op_name: str, device: str) -> list:
        return [
            {"name": _name(e["function"]), "priority": e["priority"], "device": e["device"]}
            for e in cls._candidates(op_name, device)
        ]

    @classmethod
    def get_status(cls, device: str) -> Dict[str, str]:
        """op name -> ``__name__`` of the implementation that would run on ``device`` ("none" if absent)."""
        status = {}
        for op_name in list(cls._kernels):
            fn = cls.get_best(op_name, device)
            status[op_name] = _name(fn) if fn is not None else "none"
        return status


def _name(fn: Any) -> str:
    return getattr(fn, "__name__", repr(fn))


registry = KernelRegistry()
