"""B200-native kernels behind the reference's ``src/kernels`` registry API.

Drop-in for ``/root/reference/src/kernels/__init__.py``: same module-level names
(``verify_prefix``, ``kv_append``, ``get_verify_prefix``, ``get_kv_append``, ``get_kernel_info``,
``log_kernel_status``, ``registry``), so ``src/specdec`` dispatches here with no caller changes
when this directory precedes the reference's ``src`` on ``sys.path``.

What is registered (all device "cuda", priority 100, function names containing "cuda" so that
``get_kernel_info`` reports the "cuda" backend -- reference: __init__.py:129-147):

  verify_prefix        -> cuda_ops.verify_prefix_cuda          (reference op)
  kv_append            -> cuda_ops.kv_append_cuda              (reference op)
  verify_prefix_bonus  -> cuda_ops.verify_prefix_bonus_cuda    (K+1 rows, bonus token, packed record)
  verify_step          -> cuda_ops.verify_step_cuda            (softmax-mode acceptance)
  kv_append_with_mask  -> cuda_ops.kv_append_with_mask_cuda    (reference.py:96-159 semantics)
  verify_attention     -> cuda_ops.verify_attention_cuda       ((K+1)-query decode attention)
  sample_bonus         -> cuda_ops.sample_bonus_cuda           (batched bonus-token sampler, pipeline.py:48-147)

There is no Triton entry, no PyTorch entry and no CPU entry: on a CPU-only host ``get_best``
returns None for every op (the reference's callers already handle ``None``:
policies.py:129, hf_wrappers.py:930-931), and a CUDA host whose library fails to build or load
raises at import.  The PyTorch reference implementation lives only under ``oracle/`` as the test
oracle.
"""

from __future__ import annotations

import logging
from typing import Optional

import torch

from . import cuda_ops
from .build import load_kernels
from .registry import KernelRegistry, registry

logger = logging.getLogger(__name__)

_OPS = {
    "verify_prefix": cuda_ops.verify_prefix_cuda,
    "kv_append": cuda_ops.kv_append_cuda,
    "verify_prefix_bonus": cuda_ops.verify_prefix_bonus_cuda,
    "verify_step": cuda_ops.verify_step_cuda,
    "kv_append_with_mask": cuda_ops.kv_append_with_mask_cuda,
    "verify_attention": cuda_ops.verify_attention_cuda,
    "sample_bonus": cuda_ops.sample_bonus_cuda,
}

# Build (nvcc, sm_100a) if stale and dlopen the C-ABI library.  Errors propagate.
_lib = load_kernels()


def _register_kernels() -> None:
    for op_name, fn in _OPS.items():
        registry.register(op_name, fn, priority=100, device="cuda")
    cuda_ops._register_torch_ops()


_register_kernels()


def _default_device() -> str:
    if torch.cuda.is_available():
        return "cuda"
    if torch.backends.mps.is_available():
        return "mps"
    return "cpu"


def get_verify_prefix(device: Optional[str] = None):
    """reference: src/kernels/__init__.py:84-94."""
    return registry.get_best("verify_prefix", device or _default_device())


def get_kv_append(device: Optional[str] = None):
    """reference: src/kernels/__init__.py:97-107."""
    return registry.get_best("kv_append", device or _default_device())


def get_kernel(op_name: str, device: Optional[str] = None):
    """Lookup for the ops this package adds (verify_prefix_bonus, verify_step, ...)."""
    return registry.get_best(op_name, device or _default_device())


# Import-time bindings, as in the reference (__init__.py:111-112).  None on a host without CUDA.
verify_prefix = get_verify_prefix()
kv_append = get_kv_append()


def _backend_of(fn_name: str) -> str:
    low = fn_name.lower()
    if fn_name in ("unknown", "none"):
        return "unknown"
    if "cuda" in low and "triton" not in low:
        return "cuda"
    if "triton" in low:
        return "triton"
    if "ref" in low:
        return "torch"
    return fn_name


def get_kernel_info():
    """Same keys as the reference (src/kernels/__init__.py:116-156), read by the pipeline, scheduler,
    CLI and sweep scripts."""
    device = _default_device()
    status = registry.get_status(device)
    return {
        "verify_backend": _backend_of(status.get("verify_prefix", "unknown")),
        "kv_append_backend": _backend_of(status.get("kv_append", "unknown")),
        "verify_available": get_verify_prefix(device) is not None,
        "kv_append_available": get_kv_append(device) is not None,
        "device": device,
        "library": str(_lib.path),
        "arch": "sm_100a",
    }


def log_kernel_status() -> None:
    info = get_kernel_info()
    logger.info("Kernel backends: verify=%s, kv_append=%s (%s)", info["verify_backend"], info["kv_append_backend"],
                info["library"])
    if not info["verify_available"] or not info["kv_append_available"]:
        logger.warning("sd_b200 kernels are CUDA-only; no kernel is available on device %s", info["device"])


log_kernel_status()

__all__ = [
    "verify_prefix",
    "kv_append",
    "get_verify_prefix",
    "get_kv_append",
    "get_kernel",
    "get_kernel_info",
    "log_kernel_status",
    "registry",
    "KernelRegistry",
]
