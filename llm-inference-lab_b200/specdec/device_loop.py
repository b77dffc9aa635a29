"""A verify loop that never leaves the device: producer -> fused verify step -> sequence bookkeeping, N steps, zero
host synchronisations until the caller asks for the result.

Replaces the per-sequence host block of the reference's batch path (src/specdec/core/pipeline.py:2958-3620): one
``policy.accept_tokens`` + ``.item()`` per sequence (:3018), ``.cpu().tolist()`` of the accepted tokens (:3288),
``torch.cat`` growth per sequence (:3572-3608), an extra full-prefix forward whenever all K drafts are accepted
(:3174-3206) and ~20 implicit syncs per step in token validation.  Here every quantity stays in device tensors:

    tokens   [B, T] int64   prompt + generated tokens          lengths [B] int64   valid tokens per sequence
    active   [B] bool       still generating                   prompt_len [B]      (generated = lengths - prompt_len)

and one step is: ``producer(loop)`` (the draft + teacher-forced target forward, GPU work that fills
``step.logits [B,K+1,V]``, ``step.draft_ids [B,K]`` and the arena's staged rows), ``validate_and_clamp_tokens`` on the
draft ids (device-side, specdec/utils/token_validation.py), ``VerifyStep.run()`` (ONE launch: verify + kv_append), and
``advance_sequences_device`` below -- the reference's accept / EOS / bonus / growth arithmetic as batched tensor ops.

The arithmetic of ``advance_sequences_device`` (per sequence, per step; restated on the CPU by
oracle/specdec_oracle.py::advance_sequence_reference, which the tests compare against):
  accepted = target_ids[:accept_len]                     (== the draft tokens on the accepted prefix, pipeline.py:3059)
  EOS among the accepted -> cut BEFORE it, sequence stops (pipeline.py:3116-3131 "Stop before adding EOS")
  bonus = target_ids[accept_len]                         (row accept_len; row K when everything was accepted --
                                                          replaces the extra forward of pipeline.py:3174-3206)
  bonus == EOS -> sequence stops, the bonus is still appended (pipeline.py:3265-3270)
  new = accepted + [bonus]; optional de-dup; truncate to max_new_tokens; append.
De-duplication (pipeline.py:3300-3328: drop the longest head of ``new``, up to 5 tokens, that repeats the tail of what
was generated) is a heuristic the reference added against repeated text from a misaligned KV cache; it changes greedy
output (a legitimate "the the" is dropped) and desynchronises tokens from KV rows, so it is OFF by default and
reproduced exactly when ``dedup=True`` (DESIGN.md section "de-dup")."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from specdec.utils.token_validation import validate_and_clamp_tokens
from specdec.verify_step import VerifyStep

DEDUP_MAX = 5  # pipeline.py:3311 min(5, ...)


def advance_sequences_device(tokens: torch.Tensor, lengths: torch.Tensor, prompt_len: torch.Tensor, active: torch.Tensor,
                             packed: torch.Tensor, eos_id: int, max_new_tokens: int, dedup: bool = False):
    """In-place update of (tokens, lengths, active) from the packed [B, K+2] verify records.  Pure tensor ops without
    data-dependent shapes: no host sync, graph-capturable, runs on whatever device the tensors live on.
    Returns n_new [B] (tokens appended this step)."""
    B, T = tokens.shape
    K = packed.shape[1] - 2
    dev = tokens.device
    acc = packed[:, 0].long()
    tgt = packed[:, 1:].long()                                     # [B, K+1]
    j = torch.arange(K + 1, device=dev)[None, :]
    is_acc = j < acc[:, None]
    eos_hit = (tgt == eos_id) & is_acc
    any_eos = eos_hit.any(1)
    first_eos = torch.where(any_eos, eos_hit.float().argmax(1), torch.full_like(acc, K + 1))
    n_acc = torch.minimum(acc, first_eos)                          # accepted tokens kept (a prefix)
    bonus = tgt.gather(1, acc.clamp(max=K)[:, None])               # [B, 1]
    new = torch.where(j < n_acc[:, None], tgt, bonus.expand(-1, K + 1))
    n_new = n_acc + 1
    still = active & ~any_eos & (bonus[:, 0] != eos_id)
    gen_len = lengths - prompt_len
    if dedup:
        skip = torch.zeros_like(acc)
        for c in range(min(DEDUP_MAX, K + 1), 0, -1):
            idx = (lengths[:, None] - c + torch.arange(c, device=dev)[None, :]).clamp(min=0, max=T - 1)
            same = (tokens.gather(1, idx) == new[:, :c]).all(1)
            ok = same & (gen_len >= c) & (n_new >= c) & (skip == 0)
            skip = torch.where(ok, torch.full_like(skip, c), skip)
        new = new.gather(1, (j + skip[:, None]).clamp(max=K))
        n_new = n_new - skip
    room = (max_new_tokens - gen_len).clamp(min=0)
    n_new = torch.minimum(n_new, room)
    n_new = torch.minimum(n_new, (T - lengths).clamp(min=0))
    n_new = torch.where(active, n_new, torch.zeros_like(n_new))
    pos = (lengths[:, None] + j).clamp(max=T - 1)
    valid = j < n_new[:, None]
    tokens.scatter_(1, pos, torch.where(valid, new, tokens.gather(1, pos)))
    lengths += n_new
    active.copy_(still & (lengths - prompt_len < max_new_tokens))
    return n_new


class DeviceDecodeLoop:
    """``steps`` speculative verify steps with no host synchronisation in between.

    producer(loop) must enqueue, on the current stream and without syncing, the work that fills
    ``loop.step.logits``, ``loop.step.draft_ids`` and (when there is an arena) ``loop.arena.new_k / new_v`` for the
    current ``loop.tokens / loop.lengths`` -- i.e. the draft model's K proposals and the target's teacher-forced pass."""

    def __init__(self, prompts: torch.Tensor, prompt_len: torch.Tensor, k_draft: int, vocab: int, max_new_tokens: int,
                 eos_id: int, producer: Callable[["DeviceDecodeLoop"], None], arena=None, dedup: bool = False,
                 logits_dtype: torch.dtype = torch.bfloat16, bonus_sampling: Optional[dict] = None) -> None:
        dev = prompts.device
        B, P = prompts.shape
        self.B, self.K, self.V = B, k_draft, vocab
        self.max_new_tokens, self.eos_id, self.dedup = max_new_tokens, eos_id, dedup
        self.producer, self.arena = producer, arena
        T = P + max_new_tokens + k_draft + 1
        self.tokens = torch.zeros((B, T), dtype=torch.int64, device=dev)
        self.tokens[:, :P] = prompts
        self.prompt_len = prompt_len.to(device=dev, dtype=torch.int64).clone()
        self.lengths = self.prompt_len.clone()
        self.active = torch.ones(B, dtype=torch.bool, device=dev)
        self.step = VerifyStep(B, k_draft, vocab, arena=arena, logits_dtype=logits_dtype, device=dev, bonus_sampling=bonus_sampling)
        # device-side counters (pipeline.py's per-run metrics: proposed / accepted / steps), read once at the end
        self.counters = torch.zeros(3, dtype=torch.int64, device=dev)  # proposed, accepted, steps with >= 1 active sequence

    def one_step(self) -> None:
        self.producer(self)
        self.step.draft_ids.copy_(validate_and_clamp_tokens(self.step.draft_ids, self.V, "draft_ids"))
        self.step.run()
        packed = self.step.packed
        if self.step.bonus_sampling is not None:  # do_sample: the sampled bonus replaces the greedy id of row accept_len
            packed = packed.clone()
            packed.scatter_(1, (self.step.accept_len.long() + 1)[:, None], self.step.bonus[:, None])
        n_active = self.active.sum()
        self.counters[0] += n_active * self.K
        self.counters[1] += torch.where(self.active, self.step.accept_len.long(), torch.zeros_like(self.lengths)).sum()
        self.counters[2] += (n_active > 0).long()
        advance_sequences_device(self.tokens, self.lengths, self.prompt_len, self.active, packed, self.eos_id,
                                 self.max_new_tokens, self.dedup)

    def run(self, steps: int) -> None:
        for _ in range(steps):
            self.one_step()

    def results(self):
        """The ONE synchronisation: generated tokens per sequence + counters."""
        toks, lens, pl = self.tokens.cpu(), self.lengths.cpu(), self.prompt_len.cpu()
        proposed, accepted, nsteps = (int(x) for x in self.counters.cpu().tolist())
        gen = [toks[b, int(pl[b]): int(lens[b])].tolist() for b in range(self.B)]
        return {"generated_tokens": gen, "proposed": proposed, "accepted": accepted, "steps": nsteps,
                "acceptance_rate": accepted / proposed if proposed else 0.0, "active": self.active.cpu().tolist()}
