"""Stream / event orchestration of the verify step -- drop-in for
``src/scheduler/speculative_scheduler.py`` (reference: speculative_scheduler.py:43-512).

Same class, constructor arguments, methods and metric keys as the reference
(``schedule_verification``, ``schedule_draft_generation``, ``apply_acceptance_policy``, ``get_metrics``,
``reset_metrics``, ``cleanup``, ``create_speculative_scheduler``): the draft runs on the current stream,
verification on ``verification_stream``, joined by a CUDA event (:162-292).

What is new is the part of the step that is arithmetic.  The reference verifies by running the target
model for K greedy steps (``base_model.generate_tokens``, :192-199) and then applies the policy to ONE
sequence on the host (:418-446).  ``schedule_batch_verify`` takes the target's teacher-forced
``[B, K+1, V]`` logits and runs the fused verify (+ kv_append + multi-GPU exchange) for the whole batch on
the verification stream -- eagerly or as a replayed CUDA graph -- without any host synchronisation;
``apply_acceptance_policy_batch`` is the batched, sync-free policy call.  Batch sharding across GPUs lives
in ``scheduler.batch_shard``.
"""

from __future__ import annotations

import logging
import os
import time
from typing import Any, Dict, Optional, Tuple

import torch

from kernels import get_kernel_info
from specdec.verify_step import VerifyStep

logger = logging.getLogger(__name__)
KERNELS_AVAILABLE = True


class SpeculativeScheduler:
    def __init__(self, device: str = "cuda", enable_multi_stream: Optional[bool] = None,
                 enable_batched_verification: bool = True) -> None:
        self.device = device
        if enable_multi_stream is None:
            enable_multi_stream = os.getenv("SPECDEC_PARALLEL_STREAMS", "1").lower() in ("1", "true", "yes")
        self.use_event_sync = os.getenv("SPECDEC_SYNC_MODE", "event").lower() == "event"
        self.enable_multi_stream = bool(enable_multi_stream) and str(device).startswith("cuda")
        self.enable_batched_verification = enable_batched_verification
        self.verification_stream = None
        self.default_stream = None
        self.verify_ready_event = None
        if self.enable_multi_stream:
            if not torch.cuda.is_available():
                raise RuntimeError("SpeculativeScheduler(device='cuda') needs a CUDA device: this path has no CPU fallback")
            self.verification_stream = torch.cuda.Stream()
            self.default_stream = torch.cuda.current_stream()
            self.verify_ready_event = torch.cuda.Event(enable_timing=False)
        self.kernels_available = KERNELS_AVAILABLE
        self.kernel_info = get_kernel_info()
        self.reset_metrics()
        self._steps: Dict[Tuple, VerifyStep] = {}

    # ---- reference API: model-driven verification / draft (duck-typed models) -------------------------
    def schedule_verification(self, base_model, draft_tokens, input_ids, temperature: float = 0.7,
                              do_sample: bool = True, **kwargs):
        """reference: :123-160.  The target's K greedy tokens, generated on the verification stream and
        joined to the current stream with an event (no device-wide synchronise)."""
        k = draft_tokens.shape[1]
        stream = self.verification_stream
        if stream is None:
            t0 = time.time()
            base_tokens, base_logits = base_model.generate_tokens(input_ids, max_new_tokens=k, temperature=1.0,
                                                                  do_sample=False, **kwargs)
            ms = (time.time() - t0) * 1000
            method = "single_stream"
        else:
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stream.wait_stream(torch.cuda.current_stream())
            start.record(stream)
            base_tokens, base_logits = base_model.generate_tokens(input_ids, max_new_tokens=k, temperature=1.0,
                                                                  do_sample=False, stream=stream, **kwargs)
            end.record(stream)
            end.wait(torch.cuda.current_stream())
            ms = None
            self._pending_timing = (start, end)
            method = "multi_stream"
        if ms is not None:
            self.metrics["verification_time_ms"] += ms
        self.metrics["total_steps"] += 1
        return base_tokens, base_logits, {"verification_time_ms": ms, "method": method, "stream_used": stream is not None,
                                          "sync_mode": "event" if self.use_event_sync else "barrier"}

    def schedule_draft_generation(self, draft_model, input_ids, k: int, temperature: float = 0.7,
                                  do_sample: bool = True, **kwargs):
        """reference: :370-416."""
        t0 = time.time()
        draft_tokens, draft_logits = draft_model.generate_tokens(input_ids, max_new_tokens=k, temperature=temperature,
                                                                 do_sample=do_sample, **kwargs)
        ms = (time.time() - t0) * 1000
        self.metrics["draft_time_ms"] += ms
        self.metrics["total_proposed"] += k
        return draft_tokens, draft_logits, {"draft_time_ms": ms, "k": k}

    def apply_acceptance_policy(self, policy, draft_tokens, base_tokens, draft_logits, base_logits):
        """reference: :418-446 (one sequence, returns a Python int)."""
        accepted_len, info = policy.accept_tokens(draft_tokens, base_tokens, draft_logits, base_logits)
        self.metrics["total_accepted"] += accepted_len
        return accepted_len, info

    # ---- batched, device-resident form --------------------------------------------------------------
    def apply_acceptance_policy_batch(self, policy, draft_tokens, base_logits, draft_logits=None) -> Dict[str, torch.Tensor]:
        """Whole batch, one launch, no host sync; metrics are NOT updated (that would need a D2H read)."""
        return policy.accept_batch(draft_tokens, base_logits=base_logits, proposed_logits=draft_logits)

    def schedule_batch_verify(self, logits: torch.Tensor, draft_ids: torch.Tensor, arena=None, exchange=None,
                              use_graph: bool = False, mode: str = "greedy", threshold: float = 0.0,
                              temperature: float = 1.0) -> VerifyStep:
        """Fused verify (+ kv_append of ``arena``'s staged rows, + peer exchange) of ``logits [B,K+1,V]`` /
        ``draft_ids [B,K]`` on the verification stream.  Returns the ``VerifyStep`` whose static output
        buffers (``accept_len``, ``bonus``, ``packed`` ...) are valid for work enqueued on the current
        stream after this call (event join, no synchronise)."""
        B, rows, V = logits.shape
        K = draft_ids.shape[1]
        key = (B, K, V, logits.dtype, draft_ids.dtype, id(arena), id(exchange), mode, threshold, temperature)
        step = self._steps.get(key)
        if step is None:
            step = VerifyStep(B, K, V, arena=arena, logits_dtype=logits.dtype, id_dtype=draft_ids.dtype,
                              device=logits.device, mode=mode, threshold=threshold, temperature=temperature,
                              exchange=exchange)
            self._steps[key] = step
        cur = torch.cuda.current_stream(logits.device)
        stream = self.verification_stream or cur
        stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            if use_graph:
                step.logits.copy_(logits, non_blocking=True)
                step.draft_ids.copy_(draft_ids, non_blocking=True)
                if step.graph is None:
                    step.capture()
                step.replay()
            else:
                step.run(logits, draft_ids)
            if self.verify_ready_event is not None:
                self.verify_ready_event.record(stream)
        if self.verify_ready_event is not None and stream is not cur:
            self.verify_ready_event.wait(cur)
        self.metrics["total_steps"] += 1
        self.metrics["total_proposed"] += B * K
        return step

    # ---- metrics (reference: :448-489) ----------------------------------------------------------------
    def get_metrics(self) -> Dict[str, Any]:
        pending = getattr(self, "_pending_timing", None)
        if pending is not None:  # read the event timing lazily: only this call synchronises
            pending[1].synchronize()
            self.metrics["verification_time_ms"] += pending[0].elapsed_time(pending[1])
            self._pending_timing = None
        total = self.metrics["verification_time_ms"] + self.metrics["draft_time_ms"]
        rate = self.metrics["total_accepted"] / self.metrics["total_proposed"] if self.metrics["total_proposed"] > 0 else 0.0
        return {**self.metrics, "total_time_ms": total, "acceptance_rate": rate, "device": self.device,
                "multi_stream_enabled": self.enable_multi_stream,
                "batched_verification_enabled": self.enable_batched_verification}

    def reset_metrics(self) -> None:
        self.metrics = {"total_proposed": 0, "total_accepted": 0, "total_steps": 0, "verification_time_ms": 0.0,
                        "draft_time_ms": 0.0, "overlap_time_ms": 0.0}

    def cleanup(self) -> None:
        self._steps = {}
        self.verification_stream = None


def create_speculative_scheduler(device: str = "cuda", enable_multi_stream: bool = True,
                                 enable_batched_verification: bool = True) -> SpeculativeScheduler:
    """reference: :492-512."""
    return SpeculativeScheduler(device=device, enable_multi_stream=enable_multi_stream,
                                enable_batched_verification=enable_batched_verification)
