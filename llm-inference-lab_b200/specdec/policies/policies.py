"""Acceptance policies on the B200 verify kernel -- drop-in for ``src/specdec/policies/policies.py``.

Same class names, constructor arguments, ``accept_tokens`` signature / return value and ``get_info`` /
``create_policy`` as the reference (policies.py:35-440), so ``from specdec.policies.policies import
create_policy`` works unchanged.  What differs is where the arithmetic runs:

* the reference materialises ``softmax([1,K,V])`` / ``topk`` / ``argmax`` with ATen and walks the K
  positions on the host (policies.py:156-180, :240-250, :300-314, :358-373), one sequence per call;
* here every policy is ONE launch of the fused verify kernel (``sd_verify_prefix`` / ``sd_verify_step``,
  include/sd_verify.h) over the whole batch: the acceptance rule is evaluated on the device in the same
  pass that reads the logits.  ``accept_batch`` is the batched, sync-free form (device tensors out);
  ``accept_tokens`` keeps the reference's per-sequence contract (returns a Python int, which costs
  the same ``.item()`` the reference pays, policies.py:134).

There is no PyTorch fallback: logits must be CUDA tensors (``SPECDEC_FORCE_PYTORCH_BACKEND`` is ignored
and CPU tensors raise).  When ``base_logits`` is None the reference compares token ids
(policies.py:181-196); that needs no kernel and is kept as plain tensor code.
"""

from __future__ import annotations

import logging
from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Tuple

import torch

from kernels import cuda_ops, get_kernel_info

logger = logging.getLogger(__name__)
KERNELS_AVAILABLE = True


class AcceptancePolicy(ABC):
    """reference: policies.py:35-73."""

    name = "abstract"

    @abstractmethod
    def accept_batch(self, proposed_tokens, base_logits=None, proposed_logits=None, **kwargs) -> Dict[str, torch.Tensor]:
        """Batched, sync-free: returns at least ``accept_len`` [B] int32 and ``accepted_mask`` [B,K] uint8
        on the logits' device."""

    def accept_tokens(
        self,
        proposed_tokens: torch.Tensor,
        base_tokens: torch.Tensor,
        proposed_logits: Optional[torch.Tensor] = None,
        base_logits: Optional[torch.Tensor] = None,
        **kwargs: Any,
    ) -> Tuple[int, Dict[str, Any]]:
        out = self.accept_batch(proposed_tokens, base_logits=base_logits, proposed_logits=proposed_logits,
                                base_tokens=base_tokens, **kwargs)
        accepted_len = int(out["accept_len"][0].item())  # "assuming batch size 1" (policies.py:134)
        info = {"policy": self.name, "accepted_len": accepted_len, "proposed_len": proposed_tokens.shape[1]}
        info.update(self._extra_info(out, accepted_len))
        return accepted_len, info

    def _extra_info(self, out, accepted_len) -> Dict[str, Any]:
        return {}

    @abstractmethod
    def get_info(self) -> Dict[str, Any]:
        ...


def _token_prefix(proposed_tokens: torch.Tensor, base_tokens: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Leading run of equal token ids (reference fallback without logits, policies.py:181-196)."""
    n = min(proposed_tokens.shape[1], base_tokens.shape[1])
    eq = (proposed_tokens[:, :n].long() == base_tokens[:, :n].long()).to(torch.int32)
    run = torch.cumprod(eq, dim=1)
    K = proposed_tokens.shape[1]
    mask = torch.zeros((proposed_tokens.shape[0], K), dtype=torch.uint8, device=proposed_tokens.device)
    mask[:, :n] = run.to(torch.uint8)
    return {"accept_len": run.sum(1).to(torch.int32), "accepted_mask": mask}


class LongestPrefixPolicy(AcceptancePolicy):
    """reference: policies.py:76-210 -- accept while argmax(base_logits) == proposed token."""

    def __init__(self) -> None:
        self.name = "longest_prefix"
        self.kernels_available = KERNELS_AVAILABLE
        self.kernel_info = get_kernel_info()
        logger.info("Using verify backend: %s", self.kernel_info.get("verify_backend", "unknown"))

    def accept_batch(self, proposed_tokens, base_logits=None, proposed_logits=None, base_tokens=None, **kwargs):
        if base_logits is None:
            if base_tokens is None:
                raise ValueError("LongestPrefixPolicy needs base_logits or base_tokens")
            return _token_prefix(proposed_tokens, base_tokens)
        K = proposed_tokens.shape[1]
        if base_logits.shape[1] > K:  # K+1 rows: the bonus token comes out of the same launch
            acc, mask, tgt, bonus, packed = cuda_ops.verify_prefix_bonus_cuda(base_logits[:, : K + 1], proposed_tokens)
            return {"accept_len": acc, "accepted_mask": mask, "target_ids": tgt, "bonus": bonus, "packed": packed}
        acc, mask = cuda_ops.verify_prefix_cuda(base_logits, proposed_tokens)
        return {"accept_len": acc, "accepted_mask": mask}

    def _extra_info(self, out, accepted_len):
        return {"verify_backend": self.kernel_info.get("verify_backend", "cuda"), "kernel_used": True}

    def get_info(self) -> Dict[str, Any]:
        return {"policy": self.name}


class ConfidenceThresholdPolicy(AcceptancePolicy):
    """reference: policies.py:213-269 -- accept while max softmax(proposed_logits) >= tau."""

    def __init__(self, tau: float = 0.5) -> None:
        self.tau = tau
        self.name = "conf_threshold"

    def accept_batch(self, proposed_tokens, base_logits=None, proposed_logits=None, base_tokens=None, **kwargs):
        if proposed_logits is None:  # reference falls back to the longest prefix (policies.py:235-238)
            return LongestPrefixPolicy().accept_batch(proposed_tokens, base_logits=base_logits, base_tokens=base_tokens)
        K = proposed_tokens.shape[1]
        return cuda_ops.verify_step_cuda(proposed_logits[:, :K], proposed_tokens, mode="confidence", threshold=self.tau)

    def _extra_info(self, out, accepted_len):
        extra: Dict[str, Any] = {"tau": self.tau}
        if "row_sumexp" in out:
            conf = 1.0 / out["row_sumexp"][0, :accepted_len]
            extra["min_confidence"] = float(conf.min()) if accepted_len > 0 else 0.0
        return extra

    def get_info(self) -> Dict[str, Any]:
        return {"policy": self.name, "tau": self.tau}


class TopKAgreementPolicy(AcceptancePolicy):
    """reference: policies.py:272-328 -- accept while the proposed token is in top-k(base_logits)."""

    def __init__(self, k: int = 5) -> None:
        self.k = k
        self.name = "topk_agree"

    def accept_batch(self, proposed_tokens, base_logits=None, proposed_logits=None, base_tokens=None, **kwargs):
        if proposed_logits is None or base_logits is None:  # policies.py:294-297
            return LongestPrefixPolicy().accept_batch(proposed_tokens, base_logits=base_logits, base_tokens=base_tokens)
        K = proposed_tokens.shape[1]
        return cuda_ops.verify_step_cuda(base_logits[:, :K], proposed_tokens, mode="topk", threshold=float(self.k))

    def _extra_info(self, out, accepted_len):
        return {"k": self.k}

    def get_info(self) -> Dict[str, Any]:
        return {"policy": self.name, "k": self.k}


class TypicalAcceptancePolicy(AcceptancePolicy):
    """reference: policies.py:331-396 -- accept while softmax(base_logits)[proposed token] >= p."""

    def __init__(self, p: float = 0.9) -> None:
        self.p = p
        self.name = "typical"

    def accept_batch(self, proposed_tokens, base_logits=None, proposed_logits=None, base_tokens=None, **kwargs):
        if proposed_logits is None or base_logits is None:  # policies.py:353-356
            return LongestPrefixPolicy().accept_batch(proposed_tokens, base_logits=base_logits, base_tokens=base_tokens)
        K = proposed_tokens.shape[1]
        return cuda_ops.verify_step_cuda(base_logits[:, :K], proposed_tokens, mode="typical", threshold=self.p)

    def _extra_info(self, out, accepted_len):
        extra: Dict[str, Any] = {"p": self.p}
        if "token_prob" in out:
            extra["min_probability"] = float(out["token_prob"][0, :accepted_len].min()) if accepted_len > 0 else 0.0
        return extra

    def get_info(self) -> Dict[str, Any]:
        return {"policy": self.name, "p": self.p}


def create_policy(policy_name: str, **kwargs: Any) -> AcceptancePolicy:
    """reference: policies.py:399-440."""
    policies = {
        "longest_prefix": LongestPrefixPolicy,
        "conf_threshold": lambda: ConfidenceThresholdPolicy(kwargs.get("tau", 0.5)),
        "topk_agree": lambda: TopKAgreementPolicy(kwargs.get("k", 5)),
        "typical": lambda: TypicalAcceptancePolicy(kwargs.get("p", 0.9)),
    }
    if policy_name not in policies:
        raise ValueError(f"Unknown policy: {policy_name}. Available: {list(policies.keys())}")
    return policies[policy_name]()
