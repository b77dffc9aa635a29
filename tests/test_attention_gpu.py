"""GPU parity: (K+1)-query causal decode attention (kernel 3) against the fp64 numpy oracle
(oracle/specdec_oracle.py::verify_attention) and a torch fp32 SDPA reference at full size.

Tolerance (BASELINE.json north_star): max-abs 1e-2 + rel 1e-3 for 16-bit inputs; the probabilities are
rounded to the input dtype before the P*V contraction (as every tensor-core attention does), so the
error is dominated by that rounding, not by the accumulation (fp32)."""
import numpy as np
import pytest
import torch

from kernels import cuda_ops
from oracle import specdec_oracle as O
from helpers import TORCH_DT

pytestmark = pytest.mark.gpu

ATOL, RTOL = 1e-2, 1e-3


def _case(B, Hq, Hkv, R, D, lens, cap, dtype, seed=0, poison=True):
    g = torch.Generator().manual_seed(seed)
    dt = TORCH_DT[dtype]
    q = torch.randn(B, Hq, R, D, generator=g).to(dt)
    k = torch.randn(B, Hkv, cap, D, generator=g).to(dt)
    v = torch.randn(B, Hkv, cap, D, generator=g).to(dt)
    if poison:  # rows past the valid range hold NaN: they must never reach the output
        for b, n in enumerate(lens):
            k[b, :, n + R:] = float("nan")
            v[b, :, n + R:] = float("nan")
    return q, k, v, torch.tensor(lens, dtype=torch.int32)


def _oracle(q, k, v, lens, scale=None):
    return O.verify_attention(q.float().numpy(), k.float().nan_to_num(0.0).numpy(), v.float().nan_to_num(0.0).numpy(),
                              lens.tolist(), scale)


def _check(got, want, what=""):
    got = got.float().cpu().numpy()
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    err = np.abs(got - want)
    bound = ATOL + RTOL * np.abs(want)
    assert (err <= bound).all(), f"{what}: max err {err.max():.4g} (bound {ATOL}+{RTOL}*|x|)"
    assert err.mean() < 2e-3, f"{what}: mean err {err.mean():.4g}"


SHAPES = [
    # name,               B, Hq, Hkv, R, D,   lens,                 cap
    ("llama3b",           3, 24, 8,   5, 128, [0, 17, 130],         160),
    ("llama1b",           3, 32, 8,   5, 64,  [1, 63, 200],         256),
    ("gpt2",              2, 12, 12,  5, 64,  [31, 32],             64),
    ("k1",                2, 24, 8,   2, 128, [5, 90],              128),
    ("k8_llama3b",        2, 24, 8,   9, 128, [40, 7],              64),
    ("k8_llama1b_mt3",    2, 32, 8,   9, 64,  [40, 77],             96),
    ("gqa8_headgroups",   2, 64, 8,   5, 128, [50, 3],              64),
    ("single_row",        1, 8,  8,   1, 64,  [0],                  16),
]


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", SHAPES, ids=[s[0] for s in SHAPES])
def test_against_oracle(shape, dtype):
    _, B, Hq, Hkv, R, D, lens, cap = shape
    q, k, v, sl = _case(B, Hq, Hkv, R, D, lens, cap, dtype)
    out = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), sl.cuda(), max_seq_len=max(lens))
    assert out.shape == q.shape and out.dtype == q.dtype
    _check(out, _oracle(q, k, v, sl), shape[0])


@pytest.mark.parametrize("config", [1, 2, 3])
@pytest.mark.parametrize("n_splits", [1, 2, 3, 8])
def test_splits_and_tile_configs(config, n_splits):
    B, Hq, Hkv, R, D, cap = 4, 24, 8, 5, 128, 700
    lens = [0, 100, 333, 690]
    q, k, v, sl = _case(B, Hq, Hkv, R, D, lens, cap, "bf16", seed=3)
    want = _oracle(q, k, v, sl)
    cuda_ops.set_attention_tuning(config)
    try:
        out = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), sl.cuda(), max_seq_len=max(lens), n_splits=n_splits)
    finally:
        cuda_ops.set_attention_tuning(0)
    _check(out, want, f"config {config} splits {n_splits}")


@pytest.mark.parametrize("config", [1, 2, 3])
def test_head_dim_64_configs(config):
    B, Hq, Hkv, R, D, cap = 3, 32, 8, 5, 64, 300
    lens = [2, 150, 290]
    q, k, v, sl = _case(B, Hq, Hkv, R, D, lens, cap, "fp16", seed=4)
    want = _oracle(q, k, v, sl)
    cuda_ops.set_attention_tuning(config)
    try:
        for ns in (0, 1, 4):
            out = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), sl.cuda(), max_seq_len=max(lens), n_splits=ns)
            _check(out, want, f"config {config} splits {ns}")
    finally:
        cuda_ops.set_attention_tuning(0)


def test_custom_scale_and_strided_arena():
    """The arena is a [layers,B,H,C,D] allocation: layer views have the contiguous strides the kernel
    needs; a scale other than 1/sqrt(D) is honoured."""
    B, Hq, Hkv, R, D, cap = 2, 12, 4, 5, 128, 96
    lens = [11, 80]
    q, k, v, sl = _case(B, Hq, Hkv, R, D, lens, cap, "bf16", seed=9)
    big_k = torch.full((3, B, Hkv, cap, D), float("nan"), dtype=torch.bfloat16, device="cuda")
    big_v = torch.full((3, B, Hkv, cap, D), float("nan"), dtype=torch.bfloat16, device="cuda")
    big_k[1].copy_(k)
    big_v[1].copy_(v)
    out = cuda_ops.verify_attention_cuda(q.cuda(), big_k[1], big_v[1], sl.cuda(), scale=0.2, max_seq_len=max(lens))
    _check(out, _oracle(q, k, v, sl, scale=0.2), "scale 0.2")


def test_causal_structure():
    """Query i must ignore the rows of queries j > i: changing K/V row seq+j changes outputs i >= j only."""
    B, Hq, Hkv, R, D, cap = 1, 24, 8, 5, 128, 64
    lens = [37]
    q, k, v, sl = _case(B, Hq, Hkv, R, D, lens, cap, "bf16", seed=11, poison=False)
    base = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), sl.cuda())
    for j in range(1, R):
        k2, v2 = k.clone(), v.clone()
        k2[0, :, lens[0] + j] += 3.0
        v2[0, :, lens[0] + j] -= 2.0
        out = cuda_ops.verify_attention_cuda(q.cuda(), k2.cuda(), v2.cuda(), sl.cuda())
        assert torch.equal(out[:, :, :j], base[:, :, :j]), f"row {j} leaked into earlier queries"
        assert not torch.equal(out[:, :, j:], base[:, :, j:])


def test_equals_k_sequential_decode_steps():
    """The equivalence the build relies on: one (K+1)-query pass == K+1 single-query
    decode steps over a cache grown one row at a time (what the reference's verify loop computes)."""
    B, Hq, Hkv, R, D, cap = 2, 24, 8, 5, 128, 128
    lens = [60, 99]
    q, k, v, sl = _case(B, Hq, Hkv, R, D, lens, cap, "bf16", seed=13, poison=False)
    qc, kc, vc = q.cuda(), k.cuda(), v.cuda()
    fused = cuda_ops.verify_attention_cuda(qc, kc, vc, sl.cuda())
    for i in range(R):
        step = cuda_ops.verify_attention_cuda(qc[:, :, i:i + 1].contiguous(), kc, vc, (sl + i).cuda())
        torch.testing.assert_close(step.float(), fused[:, :, i:i + 1].float(), atol=4e-3, rtol=1e-2)


def test_full_size_against_torch_fp32():
    """BASELINE configs[2] shape (B=64, L=1024, Llama-3.2-3B heads) against torch SDPA in fp32."""
    B, Hq, Hkv, R, D, L = 64, 24, 8, 5, 128, 1024
    cap = L + R
    g = torch.Generator(device="cuda").manual_seed(1234)
    q = torch.randn(B, Hq, R, D, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    sl = torch.full((B,), L, dtype=torch.int32, device="cuda")
    out = cuda_ops.verify_attention_cuda(q, k, v, sl, max_seq_len=L)
    mask = torch.arange(cap, device="cuda")[None, :] <= (L + torch.arange(R, device="cuda"))[:, None]
    kk = k.float().repeat_interleave(Hq // Hkv, dim=1)
    vv = v.float().repeat_interleave(Hq // Hkv, dim=1)
    with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.MATH):
        want = torch.nn.functional.scaled_dot_product_attention(q.float(), kk, vv, attn_mask=mask)
    _check(out, want.cpu().numpy(), "full size")
    # a few sequences also against the fp64 oracle
    idx = [0, 31, 63]
    want64 = O.verify_attention(q[idx].float().cpu().numpy(), k[idx].float().cpu().numpy(), v[idx].float().cpu().numpy(), [L] * 3)
    _check(out[idx], want64, "full size vs oracle")


# ---- balanced (persistent) form: one CTA per resident slot, equal spans of key tiles, last arriver finishes a cut item ----
@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", SHAPES, ids=[s[0] for s in SHAPES])
def test_balanced_against_oracle(shape, dtype):
    """Small shapes: fewer tiles than CTAs, so every tile is its own segment and an item is put together from as many
    partials as it has tiles -- the fix-up path carries the whole result."""
    _, B, Hq, Hkv, R, D, lens, cap = shape
    q, k, v, sl = _case(B, Hq, Hkv, R, D, lens, cap, dtype)
    ws = cuda_ops.BalancedAttentionWorkspace(B, Hq, Hkv, R, D, "cuda")
    want = _oracle(q, k, v, sl)
    for rep in range(3):  # the counters must be left zero by every launch
        out = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), sl.cuda(), max_seq_len=max(lens), balanced=ws)
        _check(out, want, f"{shape[0]} balanced, launch {rep}")
    assert int(ws.buf[:256].view(torch.int32).abs().sum()) == 0, "tile counters not reset"  # (the first 64 items' counters)


@pytest.mark.parametrize("geom", [(24, 8, 128), (32, 8, 64), (12, 12, 64)], ids=["llama3b", "llama1b", "gpt2"])
@pytest.mark.parametrize("ragged", [False, True], ids=["equal", "ragged"])
def test_balanced_full_size_equals_standard_form(geom, ragged):
    """BASELINE configs[2] / configs[3] geometries at B=64, L<=1024: spans of ~57 tiles cut most items once.  Against the
    standard form (same tiles, same per-tile arithmetic; only the fp32 merge order of a cut item differs) and the torch
    fp32 reference.  Replayed from a CUDA graph as well."""
    Hq, Hkv, D = geom
    B, R, L = 64, 5, 1024
    cap = L + R
    g = torch.Generator(device="cuda").manual_seed(77)
    q = torch.randn(B, Hq, R, D, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    if ragged:
        sl = torch.randint(0, L + 1, (B,), generator=torch.Generator().manual_seed(5), dtype=torch.int32).cuda()
        sl[0], sl[1], sl[B - 1] = 0, L, 3
    else:
        sl = torch.full((B,), L, dtype=torch.int32, device="cuda")
    ws = cuda_ops.BalancedAttentionWorkspace(B, Hq, Hkv, R, D, "cuda")
    std = cuda_ops.verify_attention_cuda(q, k, v, sl, max_seq_len=L, n_splits=1)
    bal = cuda_ops.verify_attention_cuda(q, k, v, sl, max_seq_len=L, balanced=ws)
    # one bf16 ulp (2^-7 relative) where the fp32 merge order of a cut item flips the final rounding
    excess = ((bal.float() - std.float()).abs() - (1e-3 + 2.0 ** -7 * std.float().abs())).max().item()
    assert excess <= 0, f"balanced vs standard form: {excess} beyond one bf16 ulp"
    mask = torch.arange(cap, device="cuda")[None, None, :] <= (sl.long()[:, None, None] + torch.arange(R, device="cuda")[None, :, None])
    kk = k.float().repeat_interleave(Hq // Hkv, dim=1)
    vv = v.float().repeat_interleave(Hq // Hkv, dim=1)
    with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.MATH):
        want = torch.nn.functional.scaled_dot_product_attention(q.float(), kk, vv, attn_mask=mask[:, None])
    _check(bal, want.cpu().numpy(), "balanced full size")
    # graph replay: three replays on the same workspace, every one complete and identical
    out = torch.empty_like(q)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        cuda_ops.verify_attention_cuda(q, k, v, sl, max_seq_len=L, balanced=ws, out=out)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(2):
            cuda_ops.verify_attention_cuda(q, k, v, sl, max_seq_len=L, balanced=ws, out=out)
    for _ in range(3):
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, bal), "graph replay of the balanced form differs from the eager launch"


def test_balanced_clamp_and_errors():
    B, Hq, Hkv, R, D, cap = 3, 8, 4, 5, 64, 40
    q, k, v, _ = _case(B, Hq, Hkv, R, D, [cap - R] * B, cap, "bf16", seed=3, poison=False)
    lens_ok = torch.tensor([cap - R, 7, cap - R], dtype=torch.int32)
    lens_bad = torch.tensor([cap - R, 7, cap + 9], dtype=torch.int32)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = cuda_ops.BalancedAttentionWorkspace(B, Hq, Hkv, R, D, "cuda")
    want = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), lens_ok.cuda(), overflow=flag, balanced=ws)
    assert int(flag) == 0
    _check(want, _oracle(q, k, v, lens_ok), "balanced at capacity")
    got = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), lens_bad.cuda(), overflow=flag, balanced=ws)
    assert int(flag) == 1 and torch.equal(got, want)
    with pytest.raises(AssertionError):  # a workspace made for another shape
        cuda_ops.verify_attention_cuda(q.cuda()[:2], k.cuda()[:2], v.cuda()[:2], lens_ok.cuda()[:2], balanced=ws)
    lib = cuda_ops._capi.lib()
    small = torch.zeros(512, dtype=torch.uint8, device="cuda")
    qc, kc, vc, lc = q.cuda(), k.cuda(), v.cuda(), lens_ok.cuda()
    rc = lib.sd_verify_attention(qc.data_ptr(), kc.data_ptr(), vc.data_ptr(), lc.data_ptr(), want.data_ptr(), small.data_ptr(),
                                 small.numel(), cuda_ops._DTYPE[torch.bfloat16], B, Hq, Hkv, R, D, Hkv * cap * D * 2,
                                 cap * D * 2, -1, cap, None, 0.0, cuda_ops.ATTN_BALANCED, None)
    assert rc != 0, "a workspace that cannot hold the partial slots must be refused"


def test_lengths_are_clamped_to_the_arena_capacity():
    """a sequence can sit AT capacity (the append kernels clamp there); attention must then neither read
    the next head's rows nor run off the allocation, and must say so."""
    B, Hq, Hkv, R, D, cap = 3, 8, 4, 5, 64, 40
    q, k, v, _ = _case(B, Hq, Hkv, R, D, [cap - R] * B, cap, "bf16", seed=3, poison=False)
    lens_ok = torch.tensor([cap - R, 7, cap - R], dtype=torch.int32)
    lens_bad = torch.tensor([cap - R, 7, cap + 9], dtype=torch.int32)       # beyond the slab
    lens_neg = torch.tensor([cap - R, -3, cap - R], dtype=torch.int32)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    want = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), lens_ok.cuda(), overflow=flag)
    assert int(flag) == 0
    _check(want, _oracle(q, k, v, lens_ok), "at capacity")
    got = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), lens_bad.cuda(), overflow=flag)
    assert int(flag) == 1, "clamp must be reported"
    assert torch.equal(got, want), "a length beyond the capacity must behave like capacity - R"
    flag.zero_()
    got = cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), lens_neg.cuda(), overflow=flag)
    assert int(flag) == 1 and torch.isfinite(got.float()).all()
    with pytest.raises(Exception):  # host-known bound beyond the capacity is refused
        cuda_ops.verify_attention_cuda(q.cuda(), k.cuda(), v.cuda(), lens_ok.cuda(), max_seq_len=cap)


def test_errors():
    q = torch.zeros(1, 4, 5, 128, dtype=torch.bfloat16, device="cuda")
    k = torch.zeros(1, 4, 32, 128, dtype=torch.bfloat16, device="cuda")
    sl = torch.zeros(1, dtype=torch.int32, device="cuda")
    with pytest.raises(AssertionError):
        cuda_ops.verify_attention_cuda(q.float(), k.float(), k.float(), sl)  # fp32 is not a supported dtype
    with pytest.raises(AssertionError):
        cuda_ops.verify_attention_cuda(q[:, :3], k, k, sl)  # 3 query heads over 4 KV heads
    with pytest.raises(RuntimeError):
        cuda_ops.verify_attention_cuda(q.cpu(), k.cpu(), k.cpu(), sl.cpu())
    q96 = torch.zeros(1, 4, 5, 96, dtype=torch.bfloat16, device="cuda")
    k96 = torch.zeros(1, 4, 32, 96, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(cuda_ops._capi.SdError):
        cuda_ops.verify_attention_cuda(q96, k96, k96, sl)  # head_dim 96: SD_ERR_UNSUPPORTED, never a fallback
