"""GPU: SpeculativeScheduler / GraphCapture drop-ins (stream + event orchestration around the fused step)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from scheduler.speculative_scheduler import create_speculative_scheduler
from specdec.policies import create_policy
from specdec.utils.graph_capture import GraphCapture
from helpers import make_verify_inputs, to_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_graph", [False, True])
def test_batch_verify_on_the_verification_stream(use_graph):
    sched = create_speculative_scheduler("cuda")
    assert sched.verification_stream is not None and sched.get_metrics()["multi_stream_enabled"]
    B, K, V = 5, 4, 32003
    for it in range(3):
        logits, ids = make_verify_inputs(B, K + 1, K, V, "bf16", p_match=0.7, seed=it, device="cuda")
        step = sched.schedule_batch_verify(logits, ids, use_graph=use_graph)
        got = step.packed.clone()  # enqueued on the current stream AFTER the event join: sees the results
        torch.cuda.synchronize()
        acc, _, tgt, _ = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
        np.testing.assert_array_equal(got[:, 0].cpu().numpy(), acc)
        np.testing.assert_array_equal(got[:, 1:].cpu().numpy(), tgt)
    m = sched.get_metrics()
    assert m["total_steps"] == 3 and m["total_proposed"] == 3 * B * K
    sched.cleanup()


@pytest.mark.parametrize("use_graph", [False, True])
def test_batch_verify_with_arena_appends_once_per_step(use_graph):
    """the first graphed step must not run twice -- capture()'s warm-up is side-effect free for
    the arena, so lengths and rows match the oracle from the first step on."""
    from specdec.cache.kv_arena import KVArena

    sched = create_speculative_scheduler("cuda")
    B, K, V, layers, H, D, L, cap = 6, 4, 8009, 3, 2, 64, 5, 64
    arena = KVArena(layers, B, H, D, cap, K + 1, torch.bfloat16, "cuda")
    g = torch.Generator().manual_seed(7)
    arena.k[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g).to(torch.bfloat16).cuda()
    arena.v[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g).to(torch.bfloat16).cuda()
    arena.set_lengths([L] * B)
    exp_k, exp_v = to_np(arena.k).copy(), to_np(arena.v).copy()
    lens = np.full(B, L, np.int32)
    for it in range(4):
        logits, ids = make_verify_inputs(B, K + 1, K, V, "bf16", p_match=0.7, seed=40 + it, device="cuda")
        arena.new_k.copy_(torch.randn(arena.new_k.shape, generator=g).to(torch.bfloat16))
        arena.new_v.copy_(torch.randn(arena.new_v.shape, generator=g).to(torch.bfloat16))
        step = sched.schedule_batch_verify(logits, ids, arena=arena, use_graph=use_graph)
        torch.cuda.synchronize()
        acc, _, _, _ = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
        nk, nv = to_np(arena.new_k), to_np(arena.new_v)
        for l in range(layers):
            new_lens = c_oracle.kv_arena_append(exp_k[l], lens, nk[l], acc + 1, True)
            c_oracle.kv_arena_append(exp_v[l], lens, nv[l], acc + 1, True)
        lens = new_lens
        np.testing.assert_array_equal(step.accept_len.cpu().numpy(), acc)
        np.testing.assert_array_equal(arena.seq_lens.cpu().numpy(), lens, err_msg=f"step {it}")
        # rows below the valid lengths are exact (the graph warm-up may have left rows BEYOND them at step 0)
        got_k, got_v = to_np(arena.k), to_np(arena.v)
        for b in range(B):
            np.testing.assert_array_equal(got_k[:, b, :, : lens[b]], exp_k[:, b, :, : lens[b]])
            np.testing.assert_array_equal(got_v[:, b, :, : lens[b]], exp_v[:, b, :, : lens[b]])
    sched.cleanup()


class _ToyModel:
    """Duck-typed model for the reference API: returns planted tokens / logits."""

    def __init__(self, tokens, logits):
        self.tokens, self.logits, self.seen_stream = tokens, logits, None

    def generate_tokens(self, input_ids, max_new_tokens, temperature, do_sample, stream=None, **kw):
        self.seen_stream = stream
        with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream()):
            return self.tokens[:, :max_new_tokens].clone(), self.logits[:, :max_new_tokens].clone()


def test_reference_api_round_trip():
    sched = create_speculative_scheduler("cuda")
    logits, ids = make_verify_inputs(1, 4, 4, 1000, "fp32", p_match=1.0, seed=3, device="cuda")
    ids[0, 2] = (ids[0, 2] + 1) % 1000  # mismatch at position 2
    base_tokens = torch.argmax(logits, -1)
    draft = _ToyModel(ids, logits)
    base = _ToyModel(base_tokens, logits)
    d_tok, d_log, dm = sched.schedule_draft_generation(draft, ids, k=4)
    b_tok, b_log, vm = sched.schedule_verification(base, d_tok, ids)
    assert vm["method"] == "multi_stream" and base.seen_stream is sched.verification_stream
    n, info = sched.apply_acceptance_policy(create_policy("longest_prefix"), d_tok, b_tok, d_log, b_log)
    assert n == 2 and info["accepted_len"] == 2
    m = sched.get_metrics()
    assert m["total_accepted"] == 2 and m["total_proposed"] == 4 and abs(m["acceptance_rate"] - 0.5) < 1e-9
    out = sched.apply_acceptance_policy_batch(create_policy("longest_prefix"), d_tok, b_log)
    assert out["accept_len"].tolist() == [2]
    sched.reset_metrics()
    assert sched.get_metrics()["total_steps"] == 0


def test_graph_capture_drop_in():
    gc = GraphCapture(enabled=True)
    assert gc.can_capture() and not gc.is_ready()
    x = torch.zeros(4, device="cuda")
    y = torch.zeros(4, device="cuda")

    def fn(inp):
        y.copy_(inp * 2 + 1)
        return y

    assert gc.capture_once(fn, x) and gc.is_ready()
    assert not gc.capture_once(fn, x)  # captured once (reference: graph_capture.py:60-61)
    x.fill_(3.0)
    out = gc.replay()
    torch.cuda.synchronize()
    assert out is y and y.tolist() == [7.0] * 4
    gc.reset()
    assert not gc.is_ready()
    with pytest.raises(RuntimeError):
        gc.replay()
    assert not GraphCapture(enabled=False).capture_once(fn, x)


def test_start_gate_holds_the_stream_until_released():
    """sd_start_gate, single GPU: work enqueued behind an armed gate does not start before release() (the host's
    'everything is enqueued' flag), and starts right after."""
    import time

    from kernels import cuda_ops

    gate = cuda_ops.StartGate("cuda:0")
    x = torch.zeros(1 << 20, device="cuda")
    torch.cuda.synchronize()
    gate.arm(hold=True)
    x.add_(1.0)
    done = torch.cuda.Event()
    done.record()
    time.sleep(0.05)
    assert not done.query(), "the gate let the stream through before release()"
    gate.release()
    t0 = time.perf_counter()
    while not done.query():
        assert time.perf_counter() - t0 < 5.0, "the gate never opened"
    assert float(x[0]) == 1.0
    gate.arm(hold=False)  # released before it is enqueued: passes immediately
    x.add_(1.0)
    torch.cuda.synchronize()
    assert float(x[0]) == 2.0 and cuda_ops.device_error() == 0


def test_start_gate_rendezvous_of_emulated_ranks():
    """Two ranks emulated in one process (one exchange buffer each, gates on two streams): neither stream passes its
    gate before BOTH have arrived and been released; three rounds exercise the gate counter."""
    import time

    from kernels import cuda_ops
    from scheduler.batch_shard import PeerExchange

    xs = PeerExchange.emulate(2, 8, 4, "cuda")
    try:
        gates = [cuda_ops.StartGate("cuda:0", x) for x in xs]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        flags = [torch.zeros(1, device="cuda") for _ in range(2)]
        for rnd in range(3):
            evs = []
            with torch.cuda.stream(streams[0]):
                gates[0].arm(hold=True)
                flags[0].add_(1.0)
                evs.append(torch.cuda.Event())
                evs[0].record()
            gates[0].release()
            time.sleep(0.02)
            assert not evs[0].query(), "rank 0 passed the gate before rank 1 arrived"
            with torch.cuda.stream(streams[1]):
                gates[1].arm(hold=True)
                flags[1].add_(1.0)
                evs.append(torch.cuda.Event())
                evs[1].record()
            time.sleep(0.02)
            assert not evs[0].query() and not evs[1].query(), "rank 1's host flag is still down"
            gates[1].release()
            t0 = time.perf_counter()
            while not (evs[0].query() and evs[1].query()):
                assert time.perf_counter() - t0 < 5.0, f"round {rnd}: the gate never opened"
            assert float(flags[0]) == rnd + 1 and float(flags[1]) == rnd + 1
        assert cuda_ops.device_error() == 0
    finally:
        for x in xs:
            x.close()
