"""The batched verify step: fused verify -> kv_append, device-resident, CUDA-graph capturable.

Replaces the per-sequence host loop of the reference's batch path
(src/specdec/core/pipeline.py:2958-3620: one ``policy.accept_tokens`` + ``.item()`` per sequence,
bonus-token selection :3140-3362, then ``realign_kv_cache`` :3636-3646) for the part that is
arithmetic: all sequences are verified by ONE kernel launch, the accepted KV rows of all layers are
appended by ONE more, and only a packed [B, K+2] int32 record ever needs to reach the host.

Static buffers (graph replay needs fixed addresses):
    logits     [B, K+1, V]   target logits of the teacher-forced pass (producer: the target model)
    draft_ids  [B, K]        draft tokens
    accept_len [B] i32, accepted_mask [B,K] u8, target_ids [B,K+1] i32, bonus [B] i32,
    packed     [B, K+2] i32  {accept_len, target_ids}  -> allgather / D2H record
"""

from __future__ import annotations

from typing import Optional

import torch

from kernels import _capi, cuda_ops
from specdec.cache.kv_arena import KVArena


class VerifyStep:
    def __init__(
        self,
        batch: int,
        k_draft: int,
        vocab: int,
        arena: Optional[KVArena] = None,
        logits_dtype: torch.dtype = torch.bfloat16,
        id_dtype: torch.dtype = torch.int64,
        device: str | torch.device = "cuda",
        mode: str = "greedy",
        threshold: float = 0.0,
        temperature: float = 1.0,
        exchange=None,
        bonus_sampling: Optional[dict] = None,
    ) -> None:
        self.B, self.K, self.V = batch, k_draft, vocab
        self.device = torch.device(device)
        self.arena = arena
        self.mode, self.threshold, self.temperature = mode, threshold, temperature
        self.rows_ready_early = False
        # per-sequence completion flags: verify raises flag b when sequence b is final, the append that follows
        # starts on b right away (overlaps the verify tail) and re-zeroes the flags.  Only with an arena.
        self.seq_ready = torch.zeros(batch, dtype=torch.int32, device=self.device)
        self.overlap_append = True
        self.raise_flags: Optional[bool] = None  # None = whenever this object also issues the append
        # greedy verify + append in ONE launch (sd_verify_append) when this object owns the arena
        self.fused_launch = True
        # scheduler.batch_shard.PeerExchange: the records also go to every peer GPU.  They are confirmed and
        # gathered by the append launch when there is an arena, else by a wait kernel; set
        # ``exchange_wait = False`` when the caller issues that step itself (bench.py appends separately).
        self.exchange = exchange
        self.exchange_wait = True
        # do_sample=True bonus tokens (reference: pipeline.py:48-147): dict(temperature=, top_k=, top_p=).  The
        # draw uses ``self.uniform`` [B] in [0,1) -- refill it (``step.uniform.uniform_()``) before every step;
        # ``self.bonus`` then holds the SAMPLED token of row accept_len[b] (``packed`` keeps the greedy ids).
        self.bonus_sampling = bonus_sampling
        dev = self.device
        self.uniform = torch.rand(batch, dtype=torch.float32, device=dev) if bonus_sampling else None
        self.logits = torch.zeros((batch, k_draft + 1, vocab), dtype=logits_dtype, device=dev)
        self.draft_ids = torch.zeros((batch, k_draft), dtype=id_dtype, device=dev)
        self.accept_len = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.accepted_mask = torch.zeros((batch, k_draft), dtype=torch.uint8, device=dev)
        self.target_ids = torch.zeros((batch, k_draft + 1), dtype=torch.int32, device=dev)
        self.bonus = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.packed = torch.zeros((batch, k_draft + 2), dtype=torch.int32, device=dev)
        self.token_prob = torch.zeros((batch, k_draft), dtype=torch.float32, device=dev) if mode != "greedy" else None
        self.row_max = torch.zeros((batch, k_draft + 1), dtype=torch.float32, device=dev) if mode != "greedy" else None
        self.row_sumexp = torch.zeros_like(self.row_max) if mode != "greedy" else None
        need = int(_capi.lib().sd_verify_workspace_bytes(batch, k_draft + 1))
        self.workspace = torch.zeros(need, dtype=torch.uint8, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 1 + (1 if arena is not None else 0) + (1 if bonus_sampling else 0)

    def _raise_flags(self) -> bool:
        """The flags must be consumed (and re-zeroed) by an append launch right after the verify launch; a
        caller that appends itself (bench.py) sets ``raise_flags = True`` and passes ``seq_ready`` along."""
        if self.raise_flags is not None:
            return self.raise_flags
        return self.arena is not None and self.overlap_append

    # -- eager ------------------------------------------------------------------------------------
    def run(self, logits: Optional[torch.Tensor] = None, draft_ids: Optional[torch.Tensor] = None) -> None:
        """Enqueue verify (+ append) on the current stream.  ``logits``/``draft_ids`` default to the
        static buffers; passing other tensors verifies them in place (no copy)."""
        lg = self.logits if logits is None else logits
        ids = self.draft_ids if draft_ids is None else draft_ids
        if (self.fused_launch and self.arena is not None and self.mode == "greedy" and self.overlap_append
                and self.bonus_sampling is None and cuda_ops.fused_step_available(self.arena)):
            cuda_ops.verify_append_launch(
                lg, ids, self.accept_len, self.accepted_mask, self.target_ids, self.bonus, self.packed, self.workspace,
                self.seq_ready, self.arena, zero_rejected=True, check_capacity=False,
                exchange=self.exchange if self.exchange_wait else None,
                exchange_push_in_append=False)  # finalizers push per sequence (N=8: 34.8 us vs 38.9 us per step)
            return
        cuda_ops.verify_launch(
            lg, ids, self.accept_len, self.accepted_mask, self.target_ids, self.bonus, self.packed, self.workspace,
            mode=cuda_ops.ACCEPT_MODES[self.mode], threshold=self.threshold, inv_temperature=1.0 / self.temperature,
            token_prob=self.token_prob, row_max=self.row_max, row_sumexp=self.row_sumexp,
            stats=self.mode != "greedy",
            exchange=self.exchange.desc_ptr if self.exchange is not None else None,
            seq_ready=self.seq_ready if self._raise_flags() else None,
        )
        if self.bonus_sampling is not None:
            cuda_ops.sample_bonus_cuda(lg, row_index=self.accept_len, uniform=self.uniform, out=self.bonus,
                                       temperature=self.bonus_sampling.get("temperature", 1.0),
                                       top_k=self.bonus_sampling.get("top_k"), top_p=self.bonus_sampling.get("top_p"))
        if self.arena is not None:
            # keep accept_len drafts + the token they were conditioned on (row 0 of the pass)
            # (the staged K/V rows come from the target forward, i.e. from before the verify launch, so
            # rows_ready_early would be legal here; measured on B200 it buys nothing -- the verify CTAs
            # hold the registers until the very end -- and it reads the rejected rows too)
            self.arena.append(self.accept_len, keep_bias=1, zero_rejected=True, check_capacity=False,
                              rows_ready_early=self.rows_ready_early,
                              exchange=self.exchange if self.exchange_wait else None,
                              seq_ready=self.seq_ready if self._raise_flags() else None)
        elif self.exchange is not None and self.exchange_wait:
            self.exchange.wait()

    # -- CUDA graph (reference: src/specdec/utils/graph_capture.py:49-99 GraphCapture, never wired in) --
    def capture(self) -> None:
        """Record the step into a CUDA graph.  A warm-up run (module load, kernel attributes) precedes the capture;
        it is made side-effect free for the KV arena: lengths, overflow flag and the host-side length bound are
        restored afterwards (the rows it wrote lie beyond the restored lengths and are overwritten by the first
        real step).  With a multi-GPU exchange the warm-up is a real exchange step on every rank (collective)."""
        torch.cuda.synchronize(self.device)
        saved = None
        if self.arena is not None:
            saved = (self.arena.seq_lens.clone(), self.arena.overflow.clone(), self.arena._max_len_host)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.run()  # warm-up outside capture
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if saved is not None:
            self.arena.seq_lens.copy_(saved[0])
            self.arena.overflow.copy_(saved[1])
            self.arena._max_len_host = saved[2]
            torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run()
        if saved is not None:
            self.arena._max_len_host = saved[2]  # (the capture pass advanced the host bound without running)
        self.graph = g

    def replay(self) -> None:
        if self.graph is None:
            raise RuntimeError("capture() first")
        self.graph.replay()
