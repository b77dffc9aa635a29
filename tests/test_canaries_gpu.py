"""GPU: out-of-bounds canaries.  compute-sanitizer is closed on this pool, so every output of the kernels is
placed inside a larger sentinel-filled allocation here and the sentinels around it are checked after the
launch (inputs are views with sentinel NaN padding on both sides, so an out-of-range READ would surface as a
wrong argmax / NaN as well).  The KV tests compare whole arenas bit-exactly, which covers the append kernel."""
import numpy as np
import pytest
import torch

from kernels import cuda_ops
from oracle import c_oracle
from helpers import make_verify_inputs, to_np

pytestmark = pytest.mark.gpu
PAD = 64


def _guarded(shape, dtype, sentinel):
    n = int(np.prod(shape))
    buf = torch.full((n + 2 * PAD,), sentinel, dtype=dtype, device="cuda")
    return buf, buf[PAD:PAD + n].view(shape)


def _intact(buf, n, sentinel):
    return bool((buf[:PAD] == sentinel).all()) and bool((buf[PAD + n:] == sentinel).all())


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(3, 4, 50257), (5, 2, 4099), (2, 8, 128256)])
def test_verify_outputs_stay_in_bounds(variant, shape):
    B, K, V = shape
    cuda_ops.set_verify_variant(variant, 0)
    try:
        logits, ids = make_verify_inputs(B, K + 1, K, V, "bf16", p_match=0.6, seed=V, device="cuda")
        # logits as a view into a NaN-padded buffer: a read outside [B, K+1, V] would win every argmax (NaN is maximal)
        lbuf, lview = _guarded((B, K + 1, V), torch.bfloat16, float("nan"))
        lview.copy_(logits)
        outs = {"accept_len": ((B,), torch.int32), "mask": ((B, K), torch.uint8), "tgt": ((B, K + 1), torch.int32),
                "bonus": ((B,), torch.int32), "packed": ((B, K + 2), torch.int32)}
        bufs = {k: _guarded(s, d, 113) for k, (s, d) in outs.items()}
        cuda_ops.verify_launch(lview, ids, bufs["accept_len"][1], bufs["mask"][1], bufs["tgt"][1], bufs["bonus"][1], bufs["packed"][1])
        torch.cuda.synchronize()
        for k, (s, _) in outs.items():
            assert _intact(bufs[k][0], int(np.prod(s)), 113), f"{k}: sentinel overwritten"
        acc, mask, tgt, bonus = c_oracle.verify(to_np(logits), "bf16", to_np(ids))
        np.testing.assert_array_equal(bufs["accept_len"][1].cpu().numpy(), acc)
        np.testing.assert_array_equal(bufs["tgt"][1].cpu().numpy(), tgt)
        np.testing.assert_array_equal(bufs["bonus"][1].cpu().numpy(), bonus)
    finally:
        cuda_ops.set_verify_variant()


def test_attention_and_sampler_outputs_stay_in_bounds():
    B, Hq, Hkv, R, D, cap = 3, 24, 8, 5, 128, 80
    g = torch.Generator().manual_seed(0)
    q = torch.randn(B, Hq, R, D, generator=g).to(torch.bfloat16).cuda()
    kbuf, k = _guarded((B, Hkv, cap, D), torch.bfloat16, float("nan"))
    vbuf, v = _guarded((B, Hkv, cap, D), torch.bfloat16, float("nan"))
    k.copy_(torch.randn(B, Hkv, cap, D, generator=g).to(torch.bfloat16))
    v.copy_(torch.randn(B, Hkv, cap, D, generator=g).to(torch.bfloat16))
    obuf, out = _guarded((B, Hq, R, D), torch.bfloat16, 7.0)
    lens = torch.tensor([0, 40, cap - R], dtype=torch.int32, device="cuda")  # the last sequence fills the arena exactly
    cuda_ops.verify_attention_cuda(q, k, v, lens, out=out, max_seq_len=cap - R)
    torch.cuda.synchronize()
    assert _intact(obuf, out.numel(), 7.0) and bool(torch.isfinite(out.float()).all())

    V = 50257
    lbuf, logits = _guarded((4, 3, V), torch.bfloat16, float("nan"))
    logits.copy_((torch.randn(4, 3, V, generator=g) * 2).to(torch.bfloat16))
    tbuf, tok = _guarded((4,), torch.int32, 113)
    cbuf, cnt = _guarded((4,), torch.int32, 113)
    cuda_ops.sample_bonus_cuda(logits, row_index=torch.tensor([0, 2, 1, 2], dtype=torch.int32, device="cuda"),
                               uniform=torch.tensor([0.1, 0.5, 0.9, 0.999]), temperature=0.7, top_k=50, top_p=0.9, out=tok, kept_count=cnt)
    torch.cuda.synchronize()
    assert _intact(tbuf, 4, 113) and _intact(cbuf, 4, 113)
    assert bool(((tok >= 0) & (tok < V)).all()) and bool(((cnt >= 1) & (cnt <= 50)).all())
