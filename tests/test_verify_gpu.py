"""GPU parity: the fused verify kernel (through the C ABI) against the oracle and the golden vectors.
Bit-exact: accept lengths, masks, argmax tokens, bonus tokens."""
import numpy as np
import pytest
import torch

import kernels
from kernels import cuda_ops
from oracle import c_oracle
from oracle import specdec_oracle as O
from helpers import from_np, make_verify_inputs, np_dtype_name, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[0, 1, 2, 3], ids=["tiled", "persistent", "tma_ring", "bulk_chunk"])
def verify_variant(request):
    """Every test runs under the four kernel variants (include/sd_verify.h: sd_verify_set_variant)."""
    cuda_ops.set_verify_variant(request.param, 0)
    yield request.param
    cuda_ops.set_verify_variant()


def _check_bonus(logits, ids):
    acc, mask, tgt, bonus, packed = cuda_ops.verify_prefix_bonus_cuda(logits, ids)
    torch.cuda.synchronize()
    dt = np_dtype_name(logits)
    lg = to_np(logits)
    e_acc, e_mask, e_tgt, e_bonus = c_oracle.verify(lg, dt, to_np(ids))
    np.testing.assert_array_equal(acc.cpu().numpy(), e_acc)
    np.testing.assert_array_equal(mask.cpu().numpy(), e_mask)
    np.testing.assert_array_equal(tgt.cpu().numpy(), e_tgt)
    np.testing.assert_array_equal(bonus.cpu().numpy(), e_bonus)
    np.testing.assert_array_equal(packed.cpu().numpy(), O.pack_step_result(e_acc, e_tgt))
    assert acc.dtype == torch.int32 and mask.dtype == torch.uint8 and acc.device == logits.device


def test_registry_selects_cuda_backend():
    info = kernels.get_kernel_info()
    assert info["verify_backend"] == "cuda" and info["kv_append_backend"] == "cuda"
    assert kernels.verify_prefix is cuda_ops.verify_prefix_cuda
    assert kernels.get_verify_prefix("cuda") is cuda_ops.verify_prefix_cuda


def test_golden_vectors_from_reference(golden_dir):
    """Every committed reference vector (tests/golden/verify.npz) reproduced bit-exactly."""
    z = np.load(golden_dir / "verify.npz")
    n = 0
    for entry in z["__cases__"]:
        name, dtype = str(entry).split("|")
        logits = from_np(z[f"{name}.logits"], dtype).cuda()
        ids = torch.from_numpy(z[f"{name}.ids"]).cuda()
        acc, mask = kernels.verify_prefix(logits, ids)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(acc.cpu().numpy(), z[f"{name}.accept_len"], err_msg=name)
        np.testing.assert_array_equal(mask.cpu().numpy(), z[f"{name}.mask"], err_msg=name)
        if ids.shape[1] >= 2:  # K+1-row form on the same data: all argmax tokens
            _, _, tgt, _, _ = cuda_ops.verify_prefix_bonus_cuda(logits, ids[:, :-1])
            np.testing.assert_array_equal(tgt.cpu().numpy(), z[f"{name}.argmax"], err_msg=name)
        n += 1
    assert n >= 25


def test_reference_kat_planted_matches():
    """reference: tests/test_kernels_verify.py:16-41."""
    B, K, V = 2, 3, 1000
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(B, K, V, generator=g)
    ids = torch.randint(0, V - 1, (B, K), generator=g)
    logits[0, 0, ids[0, 0]] = 10.0
    logits[0, 1, ids[0, 1]] = 10.0
    logits[0, 2, ids[0, 2] + 1] = 10.0
    logits[1, 0, ids[1, 0] + 1] = 10.0
    acc, mask = kernels.verify_prefix(logits.cuda(), ids.cuda())
    assert acc.tolist() == [2, 0]
    assert mask.tolist() == [[1, 1, 0], [0, 0, 0]]


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
@pytest.mark.parametrize("shape", [(1, 1, 100), (1, 2, 1000), (2, 3, 5000), (4, 4, 10000), (3, 5, 50257), (8, 8, 32768),
                                   (2, 4, 128256), (5, 1, 7), (3, 2, 8), (2, 3, 9), (1, 4, 4099)])
def test_vs_oracle_shapes(dtype, shape):
    B, K, V = shape
    logits, ids = make_verify_inputs(B, K + 1, K, V, dtype, p_match=0.6, seed=B * 1000 + K * 10 + V, device="cuda")
    _check_bonus(logits, ids)
    # reference signature (K rows)
    acc, mask = cuda_ops.verify_prefix_cuda(logits[:, :K], ids)
    e_acc, e_mask = O.verify_prefix(to_np(logits[:, :K]), to_np(ids), dtype)
    np.testing.assert_array_equal(acc.cpu().numpy(), e_acc)
    np.testing.assert_array_equal(mask.cpu().numpy(), e_mask)


@pytest.mark.parametrize("p_match", [0.0, 0.3, 0.8, 1.0])
def test_acceptance_rates(p_match):
    logits, ids = make_verify_inputs(16, 5, 4, 50257, "bf16", p_match=p_match, seed=11, device="cuda")
    _check_bonus(logits, ids)
    acc, _ = cuda_ops.verify_prefix_cuda(logits[:, :4], ids)
    if p_match == 1.0:
        assert acc.tolist() == [4] * 16
    if p_match == 0.0:
        assert int(acc.sum()) <= 1  # random hits are ~impossible at V=50257


def test_int32_ids_and_noncontiguous_slices():
    logits, ids = make_verify_inputs(6, 5, 4, 1536, "bf16", p_match=0.7, seed=3, device="cuda", id_dtype=torch.int32)
    _check_bonus(logits, ids)
    # per-sequence [b:b+1] slices, as the reference pipeline passes them (pipeline.py:2960-2967)
    full_acc, _, full_tgt, full_bonus, _ = cuda_ops.verify_prefix_bonus_cuda(logits, ids)
    for b in range(6):
        acc, _, tgt, bonus, _ = cuda_ops.verify_prefix_bonus_cuda(logits[b : b + 1], ids[b : b + 1])
        assert acc.item() == full_acc[b].item() and bonus.item() == full_bonus[b].item()
        assert torch.equal(tgt[0], full_tgt[b])
    # strided rows: every other row of a [B, 2R, V] buffer, and a V-slice that breaks 16-byte alignment
    big = torch.randn(4, 10, 2001, device="cuda").to(torch.bfloat16)
    view = big[:, ::2, 1:1998]
    vids = torch.argmax(view.float(), -1)[:, :4].contiguous()
    acc, _, tgt, _, _ = cuda_ops.verify_prefix_bonus_cuda(view, vids)
    e_acc, _, e_tgt, _ = c_oracle.verify(to_np(view), "bf16", to_np(vids))
    np.testing.assert_array_equal(tgt.cpu().numpy(), e_tgt)
    np.testing.assert_array_equal(acc.cpu().numpy(), e_acc)


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_adversarial_rows(dtype):
    """Ties, all-equal, +-inf, NaN, +-0 placed at every alignment phase and across chunk borders."""
    V = 70001
    g = torch.Generator().manual_seed(99)
    rows = []
    expect = []
    base = torch.randn(V, generator=g).clamp(-3, 3)

    def add(r, e):
        rows.append(r)
        expect.append(e)

    for first, second in [(17, 3000), (2047, 2048), (16383, 16384), (7, 8), (0, V - 1), (40000, 40001), (65535, 65536)]:
        r = base.clone(); r[first] = 5.0; r[second] = 5.0; add(r, first)
    add(torch.full((V,), 1.25), 0)
    add(torch.full((V,), float("-inf")), 0)
    r = base.clone(); r[777] = float("inf"); r[5] = float("-inf"); add(r, 777)
    r = base.clone(); r[60000] = float("nan"); r[60010] = float("nan"); r[9] = float("inf"); add(r, 60000)
    r = base.clone(); r[V - 1] = float("nan"); add(r, V - 1)
    r = torch.full((V,), -3.0); r[100] = -0.0; r[200] = 0.0; add(r, 100)
    r = torch.full((V,), -3.0); r[50001] = 0.0; r[50] = -0.0; add(r, 50)
    r = torch.full((V,), -3.0); r[8] = -0.0; r[9] = 0.0; r[10] = -0.0; add(r, 8)
    r = base.clone(); r[V - 1] = 9.0; add(r, V - 1)
    r = base.clone(); r[0] = 9.0; add(r, 0)
    r = base.clone().clamp(max=1.0); add(r, int(torch.argmax(r)))
    while len(rows) % 5:
        add(base.clone(), int(torch.argmax(base)))
    from helpers import TORCH_DT

    logits = torch.stack(rows).view(-1, 5, V).to(TORCH_DT[dtype])
    # expected from torch on the rounded data (ties may change after rounding)
    e_tgt = torch.argmax(logits, dim=-1)
    ids = e_tgt[:, :4].clone()
    ids[0, 2] += 1
    acc, mask, tgt, bonus, _ = cuda_ops.verify_prefix_bonus_cuda(logits.cuda(), ids.cuda())
    np.testing.assert_array_equal(tgt.cpu().numpy(), e_tgt.numpy())
    c_acc, c_mask, c_tgt, c_bonus = c_oracle.verify(to_np(logits), dtype, to_np(ids))
    np.testing.assert_array_equal(tgt.cpu().numpy(), c_tgt)
    np.testing.assert_array_equal(acc.cpu().numpy(), c_acc)
    np.testing.assert_array_equal(bonus.cpu().numpy(), c_bonus)
    if dtype == "fp32":
        assert tgt.cpu().view(-1)[: len(expect)].tolist() == expect


def test_bf16_natural_ties_large_vocab():
    """Random bf16 logits at V=128256 tie at the maximum in a few percent of rows:
    first index must win in every one of them."""
    logits, ids = make_verify_inputs(32, 5, 4, 128256, "bf16", p_match=0.0, seed=5, device="cuda")
    acc, mask, tgt, bonus, _ = cuda_ops.verify_prefix_bonus_cuda(logits, ids)
    e = torch.argmax(logits, dim=-1)
    assert torch.equal(tgt.long(), e)
    lf = logits.float()
    ties = (lf == lf.max(-1, keepdim=True).values).sum(-1) > 1
    assert ties.any(), "expected at least one tied row in this seeded sample"


@pytest.mark.parametrize("tuning", [(128, 1), (256, 2), (256, 4), (256, 8), (512, 4), (512, 8), (256, 16), (64, 1)])
def test_every_launch_configuration_is_exact(tuning):
    logits, ids = make_verify_inputs(4, 5, 4, 50257, "bf16", p_match=0.5, seed=21, device="cuda")
    lf32, _ = make_verify_inputs(2, 3, 2, 10007, "fp32", p_match=0.5, seed=22, device="cuda")
    try:
        cuda_ops.set_verify_tuning(*tuning)
        _check_bonus(logits, ids)
        _check_bonus(lf32, torch.argmax(lf32, -1)[:, :2].contiguous())
    finally:
        cuda_ops.set_verify_tuning(0, 0)


def test_empty_and_degenerate_inputs():
    dev = "cuda"
    acc, mask = cuda_ops.verify_prefix_cuda(torch.zeros(0, 3, 16, device=dev), torch.zeros(0, 3, dtype=torch.long, device=dev))
    assert acc.shape == (0,) and mask.shape == (0, 3)
    acc, mask = cuda_ops.verify_prefix_cuda(torch.zeros(2, 0, 16, device=dev), torch.zeros(2, 0, dtype=torch.long, device=dev))
    assert acc.tolist() == [0, 0] and mask.shape == (2, 0)
    # K = 0 with one bonus row: accept 0, bonus = argmax
    lg = torch.randn(3, 1, 33, device=dev)
    acc, mask, tgt, bonus, packed = cuda_ops.verify_prefix_bonus_cuda(lg, torch.zeros(3, 0, dtype=torch.long, device=dev))
    assert acc.tolist() == [0, 0, 0] and torch.equal(bonus.long(), torch.argmax(lg[:, 0], -1))
    # V = 1
    acc, mask = cuda_ops.verify_prefix_cuda(torch.randn(2, 3, 1, device=dev), torch.zeros(2, 3, dtype=torch.long, device=dev))
    assert acc.tolist() == [3, 3]
    # out-of-range / negative ids never match
    acc, _ = cuda_ops.verify_prefix_cuda(torch.randn(1, 2, 8, device=dev), torch.tensor([[-1, 2 ** 40]], device=dev))
    assert acc.tolist() == [0]


def test_current_stream_and_repeated_launches_share_workspace():
    """Launches land on torch's current stream (the reference .cu used the legacy default stream,
    verify.cu:196) and the self-resetting workspace survives back-to-back launches."""
    logits, ids = make_verify_inputs(8, 5, 4, 32000, "bf16", p_match=0.5, seed=2, device="cuda")
    ref = cuda_ops.verify_prefix_bonus_cuda(logits, ids)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        outs = [cuda_ops.verify_prefix_bonus_cuda(logits, ids) for _ in range(20)]
    s.synchronize()
    for o in outs:
        for a, b in zip(o, ref):
            assert torch.equal(a, b)


def test_full_size_llama_shape_properties():
    """BASELINE configs[2]: B=64, K=4, V=128256 bf16 -- checked against torch.argmax on device and
    through size-independent properties (prefix monotonicity, bonus consistency)."""
    logits, ids = make_verify_inputs(64, 5, 4, 128256, "bf16", p_match=0.8, seed=1234, device="cuda")
    acc, mask, tgt, bonus, packed = cuda_ops.verify_prefix_bonus_cuda(logits, ids)
    e = torch.argmax(logits, dim=-1)
    assert torch.equal(tgt.long(), e)
    match = (e[:, :4] == ids).int()
    e_acc = torch.cumprod(match, dim=1).sum(1)
    assert torch.equal(acc.long(), e_acc)
    assert torch.equal(mask.long(), (torch.arange(4, device="cuda")[None] < acc[:, None]).long())
    assert torch.equal(bonus.long(), e[torch.arange(64, device="cuda"), acc.long()])
    assert torch.equal(packed[:, 0], acc) and torch.equal(packed[:, 1:], tgt)
    # full oracle on a bounded sample of the batch
    _check_bonus(logits[:6], ids[:6])


def test_microbench_shapes_fp32():
    """scripts/microbench_verify.py:35-49 configurations (fp32 randn, 30 % planted)."""
    for (B, K, V) in [(1, 1, 4096), (1, 4, 4096), (8, 2, 4096), (8, 4, 4096), (1, 4, 8192), (1, 4, 32768)]:
        logits, ids = make_verify_inputs(B, K, K, V, "fp32", p_match=0.3, seed=V + K, device="cuda")
        acc, mask = kernels.verify_prefix(logits, ids)
        e_acc, e_mask = O.verify_prefix(to_np(logits), to_np(ids), "fp32")
        np.testing.assert_array_equal(acc.cpu().numpy(), e_acc)
        np.testing.assert_array_equal(mask.cpu().numpy(), e_mask)


def test_torch_custom_op_entry_point():
    logits, ids = make_verify_inputs(3, 4, 4, 5000, "bf16", p_match=0.5, seed=8, device="cuda")
    acc, mask = torch.ops.sd_b200.verify_prefix(logits, ids)
    e_acc, e_mask = O.verify_prefix(to_np(logits), to_np(ids), "bf16")
    np.testing.assert_array_equal(acc.cpu().numpy(), e_acc)
    np.testing.assert_array_equal(mask.cpu().numpy(), e_mask)


def test_sharded_batch_shape_b256_k8():
    """BASELINE configs[4] per-GPU shard at G=2 (B=256) with the largest K of configs[1] (K=8), V=128256 bf16
    (591 MB of logits): argmax tokens against torch.argmax on device, accept lengths / bonus through the
    leading-run property, a bounded sample against the full oracle."""
    B, K, V = 256, 8, 128256
    logits, ids = make_verify_inputs(B, K + 1, K, V, "bf16", p_match=0.85, seed=4321, device="cuda")
    acc, mask, tgt, bonus, packed = cuda_ops.verify_prefix_bonus_cuda(logits, ids)
    e = torch.argmax(logits, dim=-1)
    assert torch.equal(tgt.long(), e)
    e_acc = torch.cumprod((e[:, :K] == ids).int(), dim=1).sum(1)
    assert torch.equal(acc.long(), e_acc) and int(acc.max()) == K
    assert torch.equal(bonus.long(), e[torch.arange(B, device="cuda"), acc.long()])
    assert torch.equal(mask.sum(1).long(), acc.long())
    _check_bonus(logits[250:], ids[250:])
