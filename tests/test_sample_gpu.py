"""GPU parity of the batched bonus-token sampler (kernel 4) against the oracle restatement of
sample_bonus_token_from_logits (src/specdec/core/pipeline.py:48-147).

What can be exact is exact: the greedy branch (argmax, first index), membership of the sampled token in
the reference's filtered set, the SIZE of that set (top-k / top-p with index-ordered ties, "keep at least
one").  The draw itself is checked against the float64 CDF of the reference's filtered distribution: the
kernel's token must own the interval that contains u, within TOL of cumulative probability (the kernel
sums exp2 in 2^-40 fixed point from fp32 terms; torch.multinomial's RNG stream is not reproducible --
SURVEY.md section 7)."""
import numpy as np
import pytest
import torch

from kernels import cuda_ops
from oracle import specdec_oracle as O
from helpers import TORCH_DT, to_np

pytestmark = pytest.mark.gpu
TOL = 2e-5


def _rows(B, V, dtype, seed, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, V, generator=g) * scale).to(TORCH_DT[dtype])


def _check_draw(logits, u, tokens, kept, temperature, top_k, top_p, dtype, exact_count=True):
    x = O.as_f32(to_np(logits), dtype)
    for b in range(x.shape[0]):
        filt = O.bonus_logits_filter(x[b], temperature, top_k, top_p)
        want_tok, cdf = O.sample_bonus(x[b], temperature, top_k, top_p, float(u[b]))
        t = int(tokens[b])
        assert np.isfinite(filt[t]), f"seq {b}: token {t} is outside the reference's filtered set"
        lo = cdf[t - 1] if t > 0 else 0.0
        assert lo - TOL <= u[b] < cdf[t] + TOL, f"seq {b}: u={u[b]} not in [{lo}, {cdf[t]}) of token {t} (oracle {want_tok})"
        if exact_count:
            assert int(kept[b]) == int(np.isfinite(filt).sum()), f"seq {b}: kept {int(kept[b])} vs {int(np.isfinite(filt).sum())}"


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_greedy_branch_is_argmax(dtype):
    logits = _rows(7, 50257, dtype, 1)
    logits[0, 123] = logits[0, 4567] = 99.0  # tie: first index wins (torch.argmax)
    logits[1, :] = 0.5                       # all equal -> 0
    got = cuda_ops.sample_bonus_cuda(logits.cuda(), uniform=None)
    want = torch.argmax(logits.float(), -1)
    want[0], want[1] = 123, 0
    assert got.dtype == torch.int32 and got.cpu().tolist() == want.tolist()


@pytest.mark.parametrize("dtype,V", [("fp32", 50257), ("bf16", 128256), ("fp16", 32000), ("fp32", 1000)])
@pytest.mark.parametrize("cfg", [(1.0, None, None), (0.7, 50, 0.9), (0.7, 50, None), (1.3, None, 0.9), (1.0, 1, None),
                                 (0.5, None, 0.5), (1.0, 5, 1e-6), (2.0, 10 ** 9, 0.99)],
                         ids=["plain", "reference_default_T0.7_k50_p0.9", "topk50", "topp0.9", "topk1", "topp0.5_T0.5",
                              "topp_tiny_keeps_one", "k_ge_V"])
def test_draw_against_oracle_cdf(dtype, V, cfg):
    temperature, top_k, top_p = cfg
    B = 6
    logits = _rows(B, V, dtype, seed=V % 97 + (top_k or 0) % 7)
    g = torch.Generator().manual_seed(5)
    u = torch.rand(B, generator=g)
    u[0], u[1] = 0.0, 0.999999
    kept = torch.zeros(B, dtype=torch.int32, device="cuda")
    tok = cuda_ops.sample_bonus_cuda(logits.cuda(), uniform=u.cuda(), temperature=temperature, top_k=top_k, top_p=top_p,
                                     kept_count=kept)
    # the nucleus boundary is a float comparison: exact set size unless a cumulative sum sits within TOL of top_p
    exact = True
    if top_p is not None:
        x = O.as_f32(to_np(logits), dtype)
        for b in range(B):
            f = O.bonus_logits_filter(x[b], temperature, top_k, None).astype(np.float64)
            pr = np.sort(np.exp(f - f.max()) / np.exp(f - f.max()).sum())[::-1]
            if (np.abs(np.cumsum(pr) - top_p) < 10 * TOL).any():
                exact = False
    _check_draw(logits, u.numpy(), tok.cpu().numpy(), kept.cpu().numpy(), temperature, top_k, top_p, dtype, exact)
    if top_k == 1 or top_p == 1e-6:
        assert tok.cpu().tolist() == torch.argmax(logits.float(), -1).tolist() and kept.cpu().tolist() == [1] * B


def test_ties_at_the_threshold_are_kept_in_index_order():
    """bf16 rows are full of equal values; the reference's stable ordering keeps the lowest indices."""
    V, k = 4096, 40
    x = torch.zeros(2, V)
    x[0, 100:160] = 2.0      # 60 tied candidates for 40 places -> indices 100..139 survive
    x[1, ::2] = 1.0          # 2048 ties
    x[1, 7] = 5.0
    kept = torch.zeros(2, dtype=torch.int32, device="cuda")
    for uu in (0.0, 0.37, 0.99):
        tok = cuda_ops.sample_bonus_cuda(x.to(torch.bfloat16).cuda(), uniform=torch.tensor([uu, uu]), top_k=k, kept_count=kept)
        assert kept.cpu().tolist() == [k, k]
        assert 100 <= int(tok[0]) < 140
        t1 = int(tok[1])
        assert t1 == 7 or (t1 % 2 == 0 and t1 < 2 * (k - 1))
    _check_draw(x.to(torch.bfloat16), np.array([0.37, 0.37], np.float32), tok.cpu().numpy() * 0 + np.array(
        [int(cuda_ops.sample_bonus_cuda(x.to(torch.bfloat16).cuda(), uniform=torch.tensor([0.37, 0.37]), top_k=k)[i]) for i in range(2)]),
        kept.cpu().numpy(), 1.0, k, None, "bf16")


def test_row_selected_on_device_and_deterministic():
    B, R, V = 5, 5, 32000
    g = torch.Generator().manual_seed(9)
    logits = (torch.randn(B, R, V, generator=g) * 2).to(torch.bfloat16).cuda()
    accept = torch.tensor([0, 4, 2, 1, 3], dtype=torch.int32, device="cuda")
    u = torch.rand(B, generator=g).cuda()
    a = cuda_ops.sample_bonus_cuda(logits, row_index=accept, uniform=u, temperature=0.7, top_k=50, top_p=0.9)
    b = cuda_ops.sample_bonus_cuda(logits, row_index=accept, uniform=u, temperature=0.7, top_k=50, top_p=0.9)
    assert torch.equal(a, b)
    rows = logits[torch.arange(B), accept.long()]
    c = cuda_ops.sample_bonus_cuda(rows, uniform=u, temperature=0.7, top_k=50, top_p=0.9)
    assert torch.equal(a, c)
    greedy = cuda_ops.sample_bonus_cuda(logits, row_index=accept)
    assert greedy.cpu().tolist() == torch.argmax(rows.float(), -1).tolist()


def test_empirical_distribution_matches_the_filtered_softmax():
    """4096 draws on one small row: frequencies follow softmax(top-k filtered logits / T)."""
    V, n = 64, 4096
    g = torch.Generator().manual_seed(11)
    row = torch.randn(V, generator=g) * 1.5
    logits = row[None].expand(n, V).contiguous().cuda()
    u = torch.rand(n, generator=g).cuda()
    tok = cuda_ops.sample_bonus_cuda(logits, uniform=u, temperature=0.8, top_k=10).cpu().numpy()
    f = O.bonus_logits_filter(row.numpy(), 0.8, 10, None).astype(np.float64)
    pr = np.exp(f - f.max())
    pr /= pr.sum()
    freq = np.bincount(tok, minlength=V) / n
    assert (freq[pr == 0] == 0).all()
    assert np.abs(freq - pr).max() < 4 * np.sqrt(pr.max() / n) + 1e-3


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
@pytest.mark.parametrize("kind", ["randn", "ties", "flat", "peaked", "sorted"])
def test_fast_topk_path_equals_general_path(dtype, kind):
    """0 < top_k <= 1024 takes the two-pass candidate path (sd_sample.cuh::topk_fast_path); it must return the token
    and the kept-set size of the general radix-select path bit for bit -- random rows, rows made of a few distinct
    values (ties at the threshold, ranked by index), constant rows (everything ties: the candidate superset does not
    fit, i.e. the fallback), one dominant logit, and a descending row (all top elements in ONE thread)."""
    for (B, V) in ((16, 128256), (7, 50257), (5, 1000), (3, 37)):
        g = torch.Generator().manual_seed(V + len(kind))
        x = torch.randn(B, 3, V, generator=g) * 3.0
        if kind == "ties":
            x = torch.round(x * 2.0) / 2.0
        elif kind == "flat":
            x = torch.zeros_like(x)
        elif kind == "peaked":
            x[:, :, 7 % V] += 40.0
        elif kind == "sorted":
            x = torch.sort(x, dim=-1, descending=True).values
        logits = x.to(TORCH_DT[dtype]).cuda()
        row_index = torch.randint(0, 3, (B,), generator=g).to(torch.int32).cuda()
        u = torch.rand(B, generator=g).cuda()
        for (T, k, p_) in ((0.7, 50, 0.9), (1.0, 1, 1.0), (1.3, 1024, 0.5), (0.7, 17, 1.0), (0.5, 5, 1e-6), (1.0, 200, 0.99)):
            if k >= V:
                continue
            res = []
            for fast in (True, False):
                cuda_ops.set_sample_fast_path(fast)
                tok = torch.full((B,), -7, dtype=torch.int32, device="cuda")
                kept = torch.full((B,), -7, dtype=torch.int32, device="cuda")
                cuda_ops.sample_bonus_cuda(logits, row_index=row_index, uniform=u, temperature=T, top_k=k, top_p=p_, out=tok, kept_count=kept)
                res.append((tok.cpu().tolist(), kept.cpu().tolist()))
            cuda_ops.set_sample_fast_path(True)
            assert res[0] == res[1], (dtype, kind, V, T, k, p_)
