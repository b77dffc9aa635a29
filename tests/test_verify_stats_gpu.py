"""GPU parity of the sampling-mode pass: probabilities within a stated tolerance, accept lengths exact
away from the threshold.  Tolerance (BASELINE.json north_star): max-abs 1e-2, rel 1e-3 -- the kernel
accumulates in fp32 from exact bf16 inputs, so it is held to rel 2e-4 / abs 1e-6 here."""
import numpy as np
import pytest
import torch

from kernels import cuda_ops
from oracle import specdec_oracle as O
from helpers import make_verify_inputs, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[0, 1, 2, 3], ids=["tiled", "persistent", "tma_ring", "bulk_chunk"])
def verify_variant(request):
    """Every test runs under the four kernel variants (include/sd_verify.h: sd_verify_set_variant)."""
    cuda_ops.set_verify_variant(request.param, 0)
    yield request.param
    cuda_ops.set_verify_variant()
RTOL, ATOL = 2e-4, 1e-6


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
@pytest.mark.parametrize("shape", [(2, 3, 1000), (4, 4, 50257), (3, 4, 128256), (1, 1, 9)])
def test_softmax_statistics(dtype, shape):
    B, K, V = shape
    logits, ids = make_verify_inputs(B, K + 1, K, V, dtype, p_match=0.5, seed=V + B, device="cuda")
    out = cuda_ops.verify_step_cuda(logits, ids, mode="typical", threshold=0.3, temperature=1.0)
    x = O.as_f32(to_np(logits), dtype)
    m = x.max(-1)
    s = np.exp(x - m[..., None]).sum(-1)
    np.testing.assert_allclose(out["row_max"].cpu().numpy(), m, rtol=1e-6, atol=0)
    np.testing.assert_allclose(out["row_sumexp"].cpu().numpy(), s, rtol=RTOL)
    e_acc, e_prob = O.accept_typical(x[:, :K], to_np(ids), 0.3)
    got_prob = out["token_prob"].cpu().numpy()
    np.testing.assert_allclose(got_prob, e_prob, rtol=RTOL, atol=ATOL)
    # accept length exact unless a probability sits within tolerance of the threshold
    near = np.abs(e_prob - 0.3) < 1e-3
    if not near.any():
        np.testing.assert_array_equal(out["accept_len"].cpu().numpy(), e_acc)
    # argmax outputs are still bit-exact
    _, _, e_tgt, _ = O.verify_prefix_bonus(to_np(logits), to_np(ids), dtype)
    np.testing.assert_array_equal(out["target_ids"].cpu().numpy(), e_tgt)
    bonus = e_tgt[np.arange(B), out["accept_len"].cpu().numpy()]
    np.testing.assert_array_equal(out["bonus"].cpu().numpy(), bonus)


def test_policy_goldens_from_reference(golden_dir):
    """Reference TypicalAcceptancePolicy / ConfidenceThresholdPolicy outputs (tests/golden/policies.npz)."""
    z = np.load(golden_dir / "policies.npz")
    for name in z["__cases__"]:
        name = str(name)
        base = torch.from_numpy(z[f"{name}.base"]).cuda()
        draft = torch.from_numpy(z[f"{name}.draft"]).cuda()
        ids = torch.from_numpy(z[f"{name}.ids"]).cuda()
        K = ids.shape[1]
        ref_probs = np.take_along_axis(z[f"{name}.base_probs"], z[f"{name}.ids"][..., None], -1)[..., 0]
        for p, want in zip(z[f"{name}.typical_p"], z[f"{name}.typical"]):
            out = cuda_ops.verify_step_cuda(base, ids, mode="typical", threshold=float(p))
            np.testing.assert_allclose(out["token_prob"].cpu().numpy(), ref_probs, rtol=RTOL, atol=ATOL)
            if not (np.abs(ref_probs - p) < 1e-4).any():
                assert out["accept_len"].item() == want, (name, p)
        for t, want in zip(z[f"{name}.conf_tau"], z[f"{name}.conf"]):
            out = cuda_ops.verify_step_cuda(draft, ids, mode="confidence", threshold=float(t))
            pmax = 1.0 / out["row_sumexp"].cpu().numpy()
            if not (np.abs(pmax - t) < 1e-4).any():
                assert out["accept_len"].item() == want, (name, t)
        out = cuda_ops.verify_step_cuda(base, ids, mode="greedy")
        assert out["accept_len"].item() == z[f"{name}.longest_prefix"][0]
        assert out["bonus"].item() == z[f"{name}.bonus_greedy"][0]


def test_temperature_scaling():
    logits, ids = make_verify_inputs(3, 5, 4, 32000, "bf16", p_match=0.5, seed=77, device="cuda")
    T = 0.7  # reference default temperature (src/specdec/core/pipeline.py defaults)
    out = cuda_ops.verify_step_cuda(logits, ids, mode="typical", threshold=0.5, temperature=T)
    x = O.as_f32(to_np(logits), "bf16") / np.float32(T)
    e_acc, e_prob = O.accept_typical(x[:, :4], to_np(ids), 0.5)
    np.testing.assert_allclose(out["token_prob"].cpu().numpy(), e_prob, rtol=5e-4, atol=ATOL)


@pytest.mark.parametrize("mode,threshold", [("typical", 0.3), ("confidence", 0.5), ("topk", 5)])
def test_opt125m_vocabulary_sampling_mode_with_partial_kv_reuse(mode, threshold):
    """BASELINE configs[3]: OPT-125M target (V = 50272, 12 layers x 12 heads x 64) + GPT-2 draft with sampling-mode
    acceptance and partial KV reuse: the softmax-ratio rules evaluated by the fused verify kernel, then kv_append of the
    accepted rows into the arena and a rollback (shorter lengths) before the next step -- accept lengths exact away from
    the threshold, KV rows and lengths bit-exact."""
    from oracle import c_oracle
    from specdec.cache.kv_arena import KVArena
    from specdec.verify_step import VerifyStep

    B, K, V, layers, H, D, L, cap = 8, 4, 50272, 12, 12, 64, 9, 96
    dev = torch.device("cuda")
    arena = KVArena(layers, B, H, D, cap, K + 1, torch.float16, dev)
    g = torch.Generator().manual_seed(11)
    arena.k[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g).to(torch.float16).cuda()
    arena.v[:, :, :, :L] = torch.randn(layers, B, H, L, D, generator=g).to(torch.float16).cuda()
    arena.set_lengths([L] * B)
    T = 0.8
    step = VerifyStep(B, K, V, arena=arena, logits_dtype=torch.float16, device=dev, mode=mode, threshold=threshold, temperature=T)
    exp_k, exp_v = to_np(arena.k).copy(), to_np(arena.v).copy()
    lens = np.full(B, L, np.int32)
    for it in range(4):
        logits, ids = make_verify_inputs(B, K + 1, K, V, "fp16", p_match=0.7, seed=300 + it, device="cuda")
        step.logits.copy_(logits)
        step.draft_ids.copy_(ids)
        arena.new_k.copy_(torch.randn(arena.new_k.shape, generator=g).to(torch.float16))
        arena.new_v.copy_(torch.randn(arena.new_v.shape, generator=g).to(torch.float16))
        step.run()
        torch.cuda.synchronize()
        x = O.as_f32(to_np(logits), "fp16") / np.float32(T)
        if mode == "typical":
            e_acc, e_stat = O.accept_typical(x[:, :K], to_np(ids), threshold)
            near = np.abs(e_stat - threshold) < 1e-3
            np.testing.assert_allclose(step.token_prob.cpu().numpy(), e_stat, rtol=5e-4, atol=1e-6)
        elif mode == "confidence":
            e_acc, e_stat = O.accept_confidence(x[:, :K], threshold)
            near = np.abs(e_stat - threshold) < 1e-3
        else:
            e_acc = O.accept_topk(x[:, :K], to_np(ids), int(threshold))
            near = np.zeros(1, dtype=bool)
        acc = step.accept_len.cpu().numpy()
        if not near.any():
            np.testing.assert_array_equal(acc, e_acc)
        # the append follows the kernel's own accept lengths: KV bytes and lengths are exact whatever the rule decided
        nk, nv = to_np(arena.new_k), to_np(arena.new_v)
        for l in range(layers):
            new_lens = c_oracle.kv_arena_append(exp_k[l], lens, nk[l], acc + 1, True)
            c_oracle.kv_arena_append(exp_v[l], lens, nv[l], acc + 1, True)
        lens = new_lens
        np.testing.assert_array_equal(arena.seq_lens.cpu().numpy(), lens)
        np.testing.assert_array_equal(to_np(arena.k), exp_k)
        np.testing.assert_array_equal(to_np(arena.v), exp_v)
        # partial KV reuse: roll every second sequence back by one row (a length update; stale rows are overwritten later)
        lens = lens - (np.arange(B) % 2).astype(np.int32)
        arena.set_lengths(lens)
