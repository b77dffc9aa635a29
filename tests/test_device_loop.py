"""Sequence bookkeeping of the device-side loop against the oracle's per-sequence restatement of the reference's
arithmetic (pipeline.py:2958-3620), on CPU tensors (the function is pure tensor ops), and -- on the GPU -- the whole
DeviceDecodeLoop with a GPU-side fake logits producer under ``torch.cuda.set_sync_debug_mode("error")``: any host
synchronisation inside the loop raises."""
import numpy as np
import pytest
import torch

from oracle import specdec_oracle as O
from specdec.device_loop import DeviceDecodeLoop, advance_sequences_device
from specdec.utils.token_validation import validate_and_clamp_tokens, validation_report


def _random_case(rng, B, K, V, eos, p_eos):
    acc = rng.integers(0, K + 1, B)
    tgt = rng.integers(0, V, (B, K + 1))
    tgt[rng.random((B, K + 1)) < p_eos] = eos
    return acc, tgt


@pytest.mark.parametrize("dedup", [False, True])
@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_advance_matches_the_reference_arithmetic(dedup, device):
    rng = np.random.default_rng(3)
    B, K, V, eos, max_new = 37, 4, 11, 7, 23          # tiny vocabulary: plenty of EOS hits and tail/head overlaps
    prompt_len = torch.tensor(rng.integers(1, 6, B))
    T = 6 + max_new + K + 1
    tokens = torch.zeros(B, T, dtype=torch.int64)
    for b in range(B):
        tokens[b, : prompt_len[b]] = torch.tensor(rng.integers(0, V, int(prompt_len[b])))
    want_tokens = tokens.clone()
    lengths, active = prompt_len.clone(), torch.ones(B, dtype=torch.bool)
    tokens, lengths, active, pl = tokens.to(device), lengths.to(device), active.to(device), prompt_len.to(device)
    gen = [[] for _ in range(B)]
    act = [True] * B
    for step in range(14):
        acc, tgt = _random_case(rng, B, K, V, eos, 0.06)
        packed = torch.tensor(np.concatenate([acc[:, None], tgt], 1), dtype=torch.int32, device=device)
        advance_sequences_device(tokens, lengths, pl, active, packed, eos, max_new, dedup=dedup)
        for b in range(B):
            act[b] = O.advance_sequence_reference(gen[b], act[b], acc[b], tgt[b], eos, max_new, dedup=dedup)
        assert active.cpu().tolist() == act, f"step {step}"
        got = [tokens[b, int(prompt_len[b]): int(lengths[b])].cpu().tolist() for b in range(B)]
        assert got == gen, f"step {step}"
    assert not any(act), "the case should drive every sequence to EOS or max_new_tokens"
    assert torch.equal(tokens.cpu()[:, :1], want_tokens[:, :1])  # prompts untouched


def test_validate_and_clamp_is_the_reference_function_without_syncs():
    ids = torch.tensor([[0, 5, 99, 100, -1, 7], [100000, 3, 2, 1, 0, -5]])
    validation_report()  # clear
    out = validate_and_clamp_tokens(ids, 100, "t")
    assert out.tolist() == ids.clamp(0, 99).tolist() and out.dtype == ids.dtype
    assert validate_and_clamp_tokens(ids[:, 1:4].clamp(0, 99), 100, "t").tolist() == ids[:, 1:4].clamp(0, 99).tolist()
    rep = validation_report("t", 100)
    assert rep["t"]["invalid"] == 4 and rep["t"]["min"] == -5 and rep["t"]["max"] == 100000 and rep["t"]["calls"] == 2
    assert validation_report("t")["t"]["invalid"] == 0  # reset
    assert validate_and_clamp_tokens(None, 10) is None and validate_and_clamp_tokens(ids[:0], 10).numel() == 0


class _FakeProducer:
    """GPU-side stand-in for draft + target models (deterministic, no syncs).  The "target model" is a hash of
    (previous token, position): next = (prev * 31 + pos * 17 + 5) % V; the draft proposes the target's token except
    where a hash of (b, pos) says "miss".  Row r of the teacher-forced logits is conditioned on the DRAFT tokens before
    it, exactly like the verify pass."""

    def __init__(self, V, K, miss_mod=4):
        self.V, self.K, self.miss_mod = V, K, miss_mod

    def target_next(self, prev, pos):
        return (prev * 31 + pos * 17 + 5) % self.V

    def __call__(self, loop):
        B, K, V = loop.B, self.K, self.V
        dev = loop.tokens.device
        b = torch.arange(B, device=dev)
        L = loop.lengths
        prev = loop.tokens.gather(1, (L - 1).clamp(min=0)[:, None])[:, 0]
        drafts, targets = [], []
        for r in range(K + 1):
            pos = L + r
            t = self.target_next(prev, pos)
            targets.append(t)
            if r < K:
                miss = ((b * 7 + pos * 3) % self.miss_mod) == 0
                d = torch.where(miss, (t + 1) % V, t)
                drafts.append(d)
                prev = d
        logits = loop.step.logits
        logits.zero_()
        logits.scatter_(2, torch.stack(targets, 1)[:, :, None], torch.full((B, K + 1, 1), 8.0, dtype=logits.dtype, device=dev))
        loop.step.draft_ids.copy_(torch.stack(drafts, 1))
        if loop.arena is not None:  # staged rows: a recognisable function of (layer, b, row)
            loop.arena.new_k.copy_((L.to(torch.float32)[None, :, None, None, None] + 0.5).expand_as(loop.arena.new_k))
            loop.arena.new_v.copy_((L.to(torch.float32)[None, :, None, None, None] - 0.5).expand_as(loop.arena.new_v))


def _oracle_loop(prompts, prompt_len, prod, K, V, eos, max_new, steps, dedup):
    B = len(prompts)
    gen, act = [[] for _ in range(B)], [True] * B
    proposed = accepted = 0
    for _ in range(steps):
        for b in range(B):
            seq = prompts[b][: prompt_len[b]] + gen[b]
            prev, L = seq[-1], len(seq)
            drafts, targets = [], []
            for r in range(K + 1):
                t = (prev * 31 + (L + r) * 17 + 5) % V
                targets.append(t)
                if r < K:
                    d = (t + 1) % V if ((b * 7 + (L + r) * 3) % prod.miss_mod) == 0 else t
                    drafts.append(d)
                    prev = d
            a = 0
            while a < K and drafts[a] == targets[a]:
                a += 1
            if act[b]:
                proposed += K
                accepted += a
            act[b] = O.advance_sequence_reference(gen[b], act[b], a, targets, eos, max_new, dedup=dedup)
    return gen, act, proposed, accepted


@pytest.mark.gpu
@pytest.mark.parametrize("dedup", [False, True])
@pytest.mark.parametrize("with_arena", [False, True])
def test_device_loop_runs_without_host_syncs_and_matches_the_oracle(dedup, with_arena):
    from specdec.cache.kv_arena import KVArena

    dev = torch.device("cuda")
    B, K, V, eos, max_new, steps, P = 9, 4, 257, 256, 40, 14, 5
    rng = np.random.default_rng(1)
    prompts = rng.integers(0, V - 1, (B, P))
    prompt_len = rng.integers(2, P + 1, B)
    prod = _FakeProducer(V, K)
    arena = None
    if with_arena:
        arena = KVArena(2, B, 2, 64, P + (K + 1) * (steps + 2), K + 1, torch.bfloat16, dev)
        arena.set_lengths(prompt_len.tolist())
    loop = DeviceDecodeLoop(torch.tensor(prompts, device=dev), torch.tensor(prompt_len, device=dev), K, V, max_new, eos, prod,
                            arena=arena, dedup=dedup)
    loop.one_step()  # warm-up step outside the checked region (library load, workspace allocation) -- part of the run
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        loop.run(steps - 1)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    res = loop.results()
    gen, act, proposed, accepted = _oracle_loop(prompts.tolist(), prompt_len.tolist(), prod, K, V, eos, max_new, steps, dedup)
    assert res["generated_tokens"] == gen
    assert res["active"] == act
    assert (res["proposed"], res["accepted"]) == (proposed, accepted)
    assert 0 < res["acceptance_rate"] < 1
    assert validation_report("draft_ids")["draft_ids"]["invalid"] == 0
    if with_arena:  # every step appended accept_len + 1 rows per sequence; rows carry the length they were produced at
        lens = arena.seq_lens.cpu().numpy()
        assert (lens >= prompt_len).all() and int(arena.overflow) == 0
        k0 = arena.k[0, :, 0, :, 0].float().cpu().numpy()
        for b in range(B):
            rows = k0[b, prompt_len[b]: lens[b]]
            assert (np.diff(rows) >= 0).all() and rows[0] == prompt_len[b] + 0.5
