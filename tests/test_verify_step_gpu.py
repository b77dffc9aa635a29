"""GPU: the batched verify step (verify -> kv_append), eager and as a replayed CUDA graph."""
import time

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import specdec_oracle as O
from kernels import cuda_ops
from specdec.cache.kv_arena import KVArena
from specdec.verify_step import VerifyStep
from helpers import make_verify_inputs, to_np

pytestmark = pytest.mark.gpu


def _setup(B=6, K=4, V=50257, layers=3, H=2, D=64, L=9, cap=80):
    dev = torch.device("cuda")
    arena = KVArena(layers, B, H, D, cap, K + 1, torch.bfloat16, dev)
    g = torch.Generator().manual_seed(1)
    arena.k[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g).to(torch.bfloat16).cuda()
    arena.v[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g).to(torch.bfloat16).cuda()
    arena.set_lengths([L] * B)
    step = VerifyStep(B, K, V, arena=arena, device=dev)
    return step, arena, g


def _oracle_step(exp_k, exp_v, lens, logits, ids, nk, nv):
    acc, mask, tgt, bonus = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
    for l in range(exp_k.shape[0]):
        new_lens = c_oracle.kv_arena_append(exp_k[l], lens, nk[l], acc + 1, True)
        c_oracle.kv_arena_append(exp_v[l], lens, nv[l], acc + 1, True)
    return acc, mask, tgt, bonus, new_lens


@pytest.mark.parametrize("use_graph", [False, True])
def test_multi_step_matches_oracle(use_graph):
    step, arena, g = _setup()
    exp_k, exp_v = to_np(arena.k).copy(), to_np(arena.v).copy()
    lens = arena.seq_lens.cpu().numpy().copy()
    if use_graph:
        step.capture()                   # warm-up leaves lengths untouched; its rows lie beyond them
        np.testing.assert_array_equal(arena.seq_lens.cpu().numpy(), lens)
        arena.k.copy_(torch.from_numpy(exp_k.view(np.int16)).view(torch.bfloat16))
        arena.v.copy_(torch.from_numpy(exp_v.view(np.int16)).view(torch.bfloat16))
    for it in range(6):
        logits, ids = make_verify_inputs(step.B, step.K + 1, step.K, step.V, "bf16", p_match=0.7, seed=50 + it)
        step.logits.copy_(logits)
        step.draft_ids.copy_(ids)
        arena.new_k.copy_(torch.randn(arena.new_k.shape, generator=g).to(torch.bfloat16))
        arena.new_v.copy_(torch.randn(arena.new_v.shape, generator=g).to(torch.bfloat16))
        step.replay() if use_graph else step.run()
        torch.cuda.synchronize()
        acc, mask, tgt, bonus, lens = _oracle_step(exp_k, exp_v, lens, logits, ids, to_np(arena.new_k), to_np(arena.new_v))
        np.testing.assert_array_equal(step.accept_len.cpu().numpy(), acc)
        np.testing.assert_array_equal(step.accepted_mask.cpu().numpy(), mask)
        np.testing.assert_array_equal(step.target_ids.cpu().numpy(), tgt)
        np.testing.assert_array_equal(step.bonus.cpu().numpy(), bonus)
        np.testing.assert_array_equal(step.packed.cpu().numpy()[:, 0], acc)
        np.testing.assert_array_equal(arena.seq_lens.cpu().numpy(), lens)
        np.testing.assert_array_equal(to_np(arena.k), exp_k)
        np.testing.assert_array_equal(to_np(arena.v), exp_v)


def test_sampling_mode_step():
    dev = torch.device("cuda")
    step = VerifyStep(4, 3, 5000, arena=None, device=dev, mode="typical", threshold=0.2, temperature=0.7)
    logits, ids = make_verify_inputs(4, 4, 3, 5000, "bf16", p_match=0.6, seed=9)
    step.logits.copy_(logits)
    step.draft_ids.copy_(ids)
    step.run()
    torch.cuda.synchronize()
    from oracle import specdec_oracle as O

    x = O.as_f32(to_np(logits), "bf16") / np.float32(0.7)
    e_acc, e_prob = O.accept_typical(x[:, :3], to_np(ids), 0.2)
    np.testing.assert_allclose(step.token_prob.cpu().numpy(), e_prob, rtol=5e-4, atol=1e-6)
    if not (np.abs(e_prob - 0.2) < 1e-3).any():
        np.testing.assert_array_equal(step.accept_len.cpu().numpy(), e_acc)


def test_advance_sequences_on_device_from_packed():
    from scheduler.batch_shard import advance_sequences

    step, arena, g = _setup(B=5, K=4, V=3000, layers=1, H=1, D=16)
    logits, ids = make_verify_inputs(5, 5, 4, 3000, "bf16", p_match=0.8, seed=4)
    step.logits.copy_(logits)
    step.draft_ids.copy_(ids)
    step.run()
    tokens = torch.zeros(5, 32, dtype=torch.int64, device="cuda")
    lengths = torch.tensor([3, 4, 5, 6, 7], device="cuda")
    new_len, n_new = advance_sequences(tokens, lengths, step.packed)
    acc, _, tgt, bonus = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
    assert n_new.cpu().tolist() == (acc + 1).tolist()
    for b in range(5):
        got = tokens[b, int(lengths[b]) : int(new_len[b])].cpu().tolist()
        assert got == tgt[b, : acc[b] + 1].tolist() and got[-1] == bonus[b]
        assert got[:-1] == to_np(ids)[b, : acc[b]].tolist()  # accepted tokens are the draft tokens


@pytest.mark.parametrize("world,global_batch", [(1, 6), (2, 8), (4, 10)])
def test_peer_exchange_emulated_ranks(world, global_batch):
    """Fused result exchange: every rank's verify kernel writes its packed records into all ranks'
    exchange buffers; wait() returns the gathered [global_batch, K+2] table on every rank.  Ranks are
    emulated inside one process (all verifies are launched before any wait, so nothing spins on work
    that has not been launched); uneven shards (10 over 4) exercise the padding; several steps exercise
    the parity double-buffering and the device-side epoch."""
    from scheduler.batch_shard import PeerExchange

    K, V = 4, 4099
    xs = PeerExchange.emulate(world, global_batch, K, "cuda")
    try:
        steps = [VerifyStep(x.shard.local_batch, K, V, exchange=x) for x in xs]
        for st in steps:
            st.exchange_wait = False  # every rank's verify must be launched before any rank waits
        for it in range(5):
            logits, ids = make_verify_inputs(global_batch, K + 1, K, V, "bf16", p_match=0.6, seed=100 + it, device="cuda")
            e_acc, _, e_tgt, _ = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
            want = O.pack_step_result(e_acc, e_tgt)
            for x, st in zip(xs, steps):
                s, e = x.shard.bounds()
                if e > s:
                    st.logits.copy_(logits[s:e])
                    st.draft_ids.copy_(ids[s:e])
                    st.run()
            gathered = [x.wait().clone() for x in xs]
            torch.cuda.synchronize()
            for r, got in enumerate(gathered):
                np.testing.assert_array_equal(got.cpu().numpy(), want, err_msg=f"rank {r} step {it}")
    finally:
        for x in xs:
            x.close()


def test_peer_exchange_gathered_by_the_append_launch():
    """Same exchange, but confirmed + gathered by the first CTA of kv_append (no wait launch): the table
    lands in exchange.out, epochs advance, KV and lengths stay bit-exact."""
    from scheduler.batch_shard import PeerExchange

    world, K, V, layers, H, D, L, cap = 2, 4, 2053, 2, 2, 64, 5, 64
    global_batch = 6
    xs = PeerExchange.emulate(world, global_batch, K, "cuda")
    try:
        arenas, steps = [], []
        for x in xs:
            n = x.shard.local_batch
            a = KVArena(layers, n, H, D, cap, K + 1, torch.bfloat16, "cuda")
            a.set_lengths([L] * n)
            arenas.append(a)
            st = VerifyStep(n, K, V, arena=None, exchange=x)  # the test issues the append itself
            st.exchange_wait = False
            steps.append(st)
        lens = [np.full(x.shard.local_batch, L, np.int32) for x in xs]
        for it in range(4):
            logits, ids = make_verify_inputs(global_batch, K + 1, K, V, "bf16", p_match=0.7, seed=300 + it, device="cuda")
            e_acc, _, e_tgt, _ = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
            want = O.pack_step_result(e_acc, e_tgt)
            for x, st in zip(xs, steps):  # every rank's verify first (pushes the records) ...
                s, e = x.shard.bounds()
                st.logits.copy_(logits[s:e])
                st.draft_ids.copy_(ids[s:e])
                st.run()
            for r, (x, st, a) in enumerate(zip(xs, steps, arenas)):  # ... then the appends, which gather
                a.append(st.accept_len, keep_bias=1, exchange=x)
                s, e = x.shard.bounds()
                lens[r] = lens[r] + e_acc[s:e] + 1
            torch.cuda.synchronize()
            for r, (x, a) in enumerate(zip(xs, arenas)):
                np.testing.assert_array_equal(x.unpad(x.out).cpu().numpy(), want, err_msg=f"rank {r} step {it}")
                np.testing.assert_array_equal(a.seq_lens.cpu().numpy(), lens[r])
    finally:
        for x in xs:
            x.close()


@pytest.mark.parametrize("use_graph", [False, True])
def test_sampled_bonus_in_the_step(use_graph):
    """do_sample=True: the bonus token of row accept_len[b] is drawn on the device in the same step
    (reference defaults: temperature 0.7, top_k 50, top_p 0.9 -- pipeline.py:405-410)."""
    B, K, V = 6, 4, 32000
    step = VerifyStep(B, K, V, bonus_sampling=dict(temperature=0.7, top_k=50, top_p=0.9))
    logits, ids = make_verify_inputs(B, K + 1, K, V, "bf16", p_match=0.6, seed=21, device="cuda")
    step.logits.copy_(logits)
    step.draft_ids.copy_(ids)
    if use_graph:
        step.capture()
    for it in range(3):
        step.uniform.copy_(torch.rand(B, generator=torch.Generator().manual_seed(it)))
        step.replay() if use_graph else step.run()
        torch.cuda.synchronize()
        acc, _, tgt, _ = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
        np.testing.assert_array_equal(step.accept_len.cpu().numpy(), acc)
        x = O.as_f32(to_np(logits), "bf16")
        u = step.uniform.cpu().numpy()
        for b in range(B):
            row = x[b, acc[b]]
            filt = O.bonus_logits_filter(row, 0.7, 50, 0.9)
            _, cdf = O.sample_bonus(row, 0.7, 50, 0.9, float(u[b]))
            t = int(step.bonus[b])
            assert np.isfinite(filt[t])
            lo = cdf[t - 1] if t > 0 else 0.0
            assert lo - 2e-5 <= u[b] < cdf[t] + 2e-5


def test_peer_exchange_with_one_step_lag():
    """lag = 1: the table gathered during step t is step t-1's (nothing at t = 0); 4 slots keep epochs apart."""
    from scheduler.batch_shard import PeerExchange

    world, global_batch, K, V = 2, 6, 4, 3001
    xs = PeerExchange.emulate(world, global_batch, K, "cuda", lag=1)
    try:
        steps = [VerifyStep(x.shard.local_batch, K, V, exchange=x) for x in xs]
        for st in steps:
            st.exchange_wait = False
        for x in xs:
            x.out.fill_(-7)
        prev = None
        for it in range(7):
            logits, ids = make_verify_inputs(global_batch, K + 1, K, V, "bf16", p_match=0.6, seed=500 + it, device="cuda")
            e_acc, _, e_tgt, _ = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
            want = O.pack_step_result(e_acc, e_tgt)
            for x, st in zip(xs, steps):
                s, e = x.shard.bounds()
                st.logits.copy_(logits[s:e])
                st.draft_ids.copy_(ids[s:e])
                st.run()
            gathered = [x.wait().clone() for x in xs]
            torch.cuda.synchronize()
            for r, got in enumerate(gathered):
                if prev is None:
                    assert (got.cpu().numpy() == -7).all()  # nothing gathered at the first step
                else:
                    np.testing.assert_array_equal(got.cpu().numpy(), prev, err_msg=f"rank {r} step {it}")
            prev = want
    finally:
        for x in xs:
            x.close()


def test_peer_exchange_pushed_and_gathered_by_the_append_launch():
    """exchange_packed: the FIRST CTA of kv_append pushes this rank's records, its LAST CTA gathers.  Ranks
    emulated in one process can only run this with lag = 1 (with lag 0 a rank's append would spin on a
    push that is launched after it); real ranks are checked by tools/check_peer_exchange.py."""
    from scheduler.batch_shard import PeerExchange

    world, K, V, layers, H, D, L, cap = 2, 4, 2053, 2, 2, 64, 5, 64
    global_batch = 6
    xs = PeerExchange.emulate(world, global_batch, K, "cuda", lag=1)
    try:
        arenas, steps = [], []
        for x in xs:
            n = x.shard.local_batch
            a = KVArena(layers, n, H, D, cap, K + 1, torch.bfloat16, "cuda")
            a.set_lengths([L] * n)
            arenas.append(a)
            steps.append(VerifyStep(n, K, V, arena=None))  # the verify launch does not carry the exchange
            x.out.fill_(-7)
        prev = None
        for it in range(5):
            logits, ids = make_verify_inputs(global_batch, K + 1, K, V, "bf16", p_match=0.7, seed=700 + it, device="cuda")
            e_acc, _, e_tgt, _ = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
            want = O.pack_step_result(e_acc, e_tgt)
            for x, st, a in zip(xs, steps, arenas):
                s, e = x.shard.bounds()
                st.logits.copy_(logits[s:e])
                st.draft_ids.copy_(ids[s:e])
                st.run()
                a.append(st.accept_len, keep_bias=1, exchange=x, exchange_packed=st.packed)
            torch.cuda.synchronize()
            for r, x in enumerate(xs):
                got = x.unpad(x.out).cpu().numpy()
                if prev is None:
                    assert (got == -7).all()
                else:
                    np.testing.assert_array_equal(got, prev, err_msg=f"rank {r} step {it}")
            prev = want
    finally:
        for x in xs:
            x.close()


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("overlap", [True, False, "r1_layout", "lag1", "lag3", "lag64"])
def test_append_overlapping_the_verify_tail(use_graph, overlap):
    """Per-sequence flags (seq_ready): the append starts on a sequence as soon as that sequence is verified.
    Same results as the serialised form, over many steps (flags re-armed every step), Llama-like row size."""
    B, K, V, layers, H, D, L, cap = 48, 4, 32000, 4, 8, 128, 7, 96
    dev = torch.device("cuda")
    arena = KVArena(layers, B, H, D, cap, K + 1, torch.bfloat16, dev)
    g = torch.Generator().manual_seed(2)
    arena.k[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g).to(torch.bfloat16).cuda()
    arena.v[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g).to(torch.bfloat16).cuda()
    arena.set_lengths([L] * B)
    step = VerifyStep(B, K, V, arena=arena, device=dev)
    # True = ONE launch, lagged append (heuristic lag); lagN = the same with N sequences of lag (64 > B: every
    # piece in the tail); r1_layout = ONE launch, append CTAs behind all verify CTAs; False = two launches
    step.overlap_append = bool(overlap)
    from kernels import cuda_ops as _ops

    if overlap == "r1_layout":
        _ops.set_fused_mode(1, 0)
    elif isinstance(overlap, str):
        _ops.set_fused_mode(0, int(overlap[3:]))
    try:
        exp_k, exp_v = to_np(arena.k).copy(), to_np(arena.v).copy()
        lens = np.full(B, L, np.int32)
        if use_graph:
            step.capture()
            arena.set_lengths([L] * B)
            arena.k.copy_(torch.from_numpy(exp_k.view(np.int16)).view(torch.bfloat16))
            arena.v.copy_(torch.from_numpy(exp_v.view(np.int16)).view(torch.bfloat16))
        for it in range(6):
            logits, ids = make_verify_inputs(B, K + 1, K, V, "bf16", p_match=0.75, seed=900 + it, device="cuda")
            step.logits.copy_(logits)
            step.draft_ids.copy_(ids)
            arena.new_k.copy_(torch.randn(arena.new_k.shape, generator=g).to(torch.bfloat16))
            arena.new_v.copy_(torch.randn(arena.new_v.shape, generator=g).to(torch.bfloat16))
            torch.cuda.synchronize()
            step.replay() if use_graph else step.run()
            torch.cuda.synchronize()
            acc, mask, tgt, bonus, lens = _oracle_step(exp_k, exp_v, lens, logits, ids, to_np(arena.new_k), to_np(arena.new_v))
            np.testing.assert_array_equal(step.accept_len.cpu().numpy(), acc)
            np.testing.assert_array_equal(arena.seq_lens.cpu().numpy(), lens)
            np.testing.assert_array_equal(to_np(arena.k), exp_k)
            np.testing.assert_array_equal(to_np(arena.v), exp_v)
            assert int(step.seq_ready.sum()) == 0  # re-armed
            assert int(arena._ticket.item()) == 0
    finally:
        _ops.set_fused_mode(0, 0)
    assert _ops.device_error() == 0


@pytest.mark.parametrize("world", [1, 2])
def test_fused_launch_with_finalizer_push_and_one_step_lag(world):
    """Fused launch, records pushed by the verify finalizers, lag = 1: the gathering append CTA must not advance
    the epoch before every finalizer of the same launch has read it (the round-1 multi-GPU stall,
    tools/repro_epoch_race.py).  Many steps at a size where finalizers are still running when the first append
    CTA is done."""
    from scheduler.batch_shard import PeerExchange

    K, V, layers, H, D, L = 4, 32000, 4, 8, 128, 7
    global_batch = 48 * world
    xs = PeerExchange.emulate(world, global_batch, K, "cuda", lag=1)
    try:
        arenas, steps = [], []
        for x in xs:
            n = x.shard.local_batch
            a = KVArena(layers, n, H, D, L + 5 * 40, K + 1, torch.bfloat16, "cuda")
            a.set_lengths([L] * n)
            a.new_k.normal_()
            a.new_v.normal_()
            arenas.append(a)
            steps.append(VerifyStep(n, K, V, arena=a, exchange=x))  # run(): ONE launch, finalizers push, append gathers
            x.out.fill_(-7)
        prev = None
        for it in range(40):
            logits, ids = make_verify_inputs(global_batch, K + 1, K, V, "bf16", p_match=0.7, seed=900 + it, device="cuda")
            for x, st, a in zip(xs, steps, arenas):
                s, e = x.shard.bounds()
                st.logits.copy_(logits[s:e])
                st.draft_ids.copy_(ids[s:e])
                if it % 2:
                    st.run()
                else:  # the same launch through the op
                    cuda_ops.verify_append_launch(st.logits, st.draft_ids, st.accept_len, st.accepted_mask, st.target_ids,
                                                  st.bonus, st.packed, st.workspace, st.seq_ready, a, exchange=x,
                                                  exchange_push_in_append=False)
            done = torch.cuda.Event()
            done.record()
            t0 = time.perf_counter()
            while not done.query():
                assert time.perf_counter() - t0 < 5.0, f"step {it} stalled"
                time.sleep(0.001)
            want = torch.cat([st.packed for st in steps]).cpu().numpy()
            for r, x in enumerate(xs):
                got = x.unpad(x.out).cpu().numpy()
                if prev is None:
                    assert (got == -7).all()
                else:
                    np.testing.assert_array_equal(got, prev, err_msg=f"rank {r} step {it}")
            prev = want
    finally:
        for x in xs:
            x.close()


@pytest.mark.parametrize("lag", [0, 5])
def test_fused_step_with_a_grid_larger_than_residency(lag):
    """B = 256, V = 128256: 9000+ CTAs against 592 resident ones, so most CTAs a waiting CTA depends on are long gone
    and the last ones are dispatched late -- the dispatch-order assumption documented in include/sd_verify.h
    (sd_verify_append) at scale.  Small lag (5 sequences < one wave) makes carrying CTAs really wait.  Results vs the
    oracle; no bounded wait may have expired."""
    B, K, V, layers, H, D, L, cap = 256, 4, 128256, 2, 8, 128, 3, 24
    dev = torch.device("cuda")
    arena = KVArena(layers, B, H, D, cap, K + 1, torch.bfloat16, dev)
    g = torch.Generator(device="cuda").manual_seed(5)
    arena.k[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g, device="cuda").to(torch.bfloat16)
    arena.v[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g, device="cuda").to(torch.bfloat16)
    arena.set_lengths([L] * B)
    step = VerifyStep(B, K, V, arena=arena, device=dev)
    cuda_ops.set_fused_mode(0, lag)
    try:
        exp_k, exp_v = to_np(arena.k).copy(), to_np(arena.v).copy()
        lens = np.full(B, L, np.int32)
        for it in range(3):
            gl = torch.Generator(device="cuda").manual_seed(40 + it)
            step.logits.copy_(torch.randn(B, K + 1, V, generator=gl, device="cuda").to(torch.bfloat16))
            ids = torch.randint(0, V, (B, K), generator=gl, device="cuda")
            plant = torch.rand(B, K, generator=gl, device="cuda") < 0.75
            bi, ki = torch.nonzero(plant, as_tuple=True)
            step.logits[bi, ki, ids[bi, ki]] = 10.0
            step.draft_ids.copy_(ids)
            arena.new_k.copy_(torch.randn(arena.new_k.shape, generator=gl, device="cuda").to(torch.bfloat16))
            arena.new_v.copy_(torch.randn(arena.new_v.shape, generator=gl, device="cuda").to(torch.bfloat16))
            step.run()
            torch.cuda.synchronize()
            acc, mask, tgt, bonus, lens = _oracle_step(exp_k, exp_v, lens, step.logits, step.draft_ids, to_np(arena.new_k), to_np(arena.new_v))
            np.testing.assert_array_equal(step.accept_len.cpu().numpy(), acc)
            np.testing.assert_array_equal(step.target_ids.cpu().numpy(), tgt)
            np.testing.assert_array_equal(arena.seq_lens.cpu().numpy(), lens)
            np.testing.assert_array_equal(to_np(arena.k), exp_k)
            np.testing.assert_array_equal(to_np(arena.v), exp_v)
    finally:
        cuda_ops.set_fused_mode(0, 0)
    assert cuda_ops.device_error() == 0
