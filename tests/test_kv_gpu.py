"""GPU parity: kv_append (functional drop-in, masked variant, all-layer arena form) -- bit-exact copies."""
import numpy as np
import pytest
import torch

import kernels
from kernels import cuda_ops
from oracle import c_oracle
from oracle import specdec_oracle as O
from specdec.cache.kv_arena import KVArena
from helpers import TORCH_DT, from_np, to_np

pytestmark = pytest.mark.gpu


def _rand(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(TORCH_DT[dtype]).cuda()


def test_golden_vectors_from_reference(golden_dir):
    z = np.load(golden_dir / "kv_append.npz")
    n = 0
    for entry in z["__cases__"]:
        name, dtype, _ = str(entry).split("|")
        bk, bv, nk, nv = (from_np(z[f"{name}.{k}"], dtype).cuda() for k in ("bk", "bv", "nk", "nv"))
        ok, ov = kernels.kv_append(bk, bv, nk, nv)
        np.testing.assert_array_equal(to_np(ok), z[f"{name}.ok"], err_msg=name)
        np.testing.assert_array_equal(to_np(ov), z[f"{name}.ov"], err_msg=name)
        mask = torch.from_numpy(z[f"{name}.mask"]).cuda()
        alen = torch.from_numpy(z[f"{name}.alen"]).cuda()
        mk, mv = cuda_ops.kv_append_with_mask_cuda(bk, bv, nk, nv, mask, alen)
        np.testing.assert_array_equal(to_np(mk), z[f"{name}.mk"], err_msg=name)
        np.testing.assert_array_equal(to_np(mv), z[f"{name}.mv"], err_msg=name)
        if f"{name}.hk" in z:  # non-prefix mask: compaction
            hk, hv = cuda_ops.kv_append_with_mask_cuda(
                bk, bv, nk, nv, torch.from_numpy(z[f"{name}.hole_mask"]).cuda(), torch.from_numpy(z[f"{name}.hole_alen"]).cuda())
            np.testing.assert_array_equal(to_np(hk), z[f"{name}.hk"], err_msg=name)
            np.testing.assert_array_equal(to_np(hv), z[f"{name}.hv"], err_msg=name)
        n += 1
    assert n == 18


def test_reference_kats():
    """reference: tests/test_kv_cache.py:15-84 and :119-186."""
    B, H, L, D, K = 1, 2, 3, 4, 2
    bk, bv = torch.randn(B, H, L, D).cuda(), torch.randn(B, H, L, D).cuda()
    nk, nv = torch.randn(B, H, K, D).cuda(), torch.randn(B, H, K, D).cuda()
    ok, ov = kernels.kv_append(bk, bv, nk, nv)
    assert ok.shape == (B, H, L + K, D)
    assert torch.equal(ok[:, :, :L], bk) and torch.equal(ov[:, :, :L], bv)
    assert torch.equal(ok[:, :, L:], nk) and torch.equal(ov[:, :, L:], nv)
    # zero acceptance keeps the base and zero-fills the tail (test_kv_cache.py:165-186)
    mk, mv = cuda_ops.kv_append_with_mask_cuda(bk, bv, nk, nv, torch.zeros(B, K, dtype=torch.uint8), torch.zeros(B, dtype=torch.int32))
    assert torch.equal(mk[:, :, :L], bk) and not mk[:, :, L:].any()
    # partial [1,1,0,0]
    K = 4
    dk = torch.arange(K).view(1, 1, K, 1).expand(B, H, K, D).float().contiguous().cuda()
    mk, mv = cuda_ops.kv_append_with_mask_cuda(torch.zeros(B, H, L, D).cuda(), torch.zeros(B, H, L, D).cuda(), dk, dk * 10,
                                               torch.tensor([[1, 1, 0, 0]], dtype=torch.uint8), torch.tensor([2], dtype=torch.int32))
    assert torch.equal(mk[:, :, L], dk[:, :, 0]) and torch.equal(mk[:, :, L + 1], dk[:, :, 1]) and not mk[:, :, L + 2 :].any()


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_repeated_appends_reuse_the_arena(dtype):
    """cache = kv_append(cache, chunk) in a loop == torch.cat of everything; only new rows move."""
    B, H, D = 3, 4, 64
    k, v = _rand((B, H, 5, D), dtype, 1), _rand((B, H, 5, D), dtype, 2)
    want_k, want_v = k.clone(), v.clone()
    ptrs = set()
    for step in range(40):
        n = 1 + step % 5
        nk, nv = _rand((B, H, n, D), dtype, 100 + step), _rand((B, H, n, D), dtype, 200 + step)
        k, v = kernels.kv_append(k, v, nk, nv)
        want_k, want_v = torch.cat([want_k, nk], 2), torch.cat([want_v, nv], 2)
        ptrs.add(k.data_ptr())
        assert torch.equal(k, want_k) and torch.equal(v, want_v)
    assert len(ptrs) <= 4, "arena should be reallocated O(log n) times, not every step"


def test_rollback_then_append():
    """slice_prefix (rollback) followed by append keeps exact contents (reference flow: kv_types.py:62-86 then
    hf_wrappers.py:1009) AND leaves the tensor it was sliced from untouched, as torch.cat would: a prefix shorter than
    the arena's tip is a stale handle, so the result lives in a fresh arena (no silent overwrite)."""
    B, H, D = 2, 2, 32
    k, v = _rand((B, H, 8, D), "bf16", 5), _rand((B, H, 8, D), "bf16", 6)
    c1, c2 = _rand((B, H, 4, D), "bf16", 7), _rand((B, H, 4, D), "bf16", 8)
    k1, v1 = kernels.kv_append(k, v, c1, c2)
    k2, v2 = kernels.kv_append(k1[:, :, :10], v1[:, :, :10], c2[:, :, :3], c1[:, :, :3])
    assert torch.equal(k2, torch.cat([k, c1[:, :, :2], c2[:, :, :3]], 2))
    assert torch.equal(v2, torch.cat([v, c2[:, :, :2], c1[:, :, :3]], 2))
    assert torch.equal(k1, torch.cat([k, c1], 2)) and torch.equal(v1, torch.cat([v, c2], 2)), "rows of the older result were overwritten"
    k3, v3 = kernels.kv_append(k2, v2, c1[:, :, :1], c2[:, :, :1])  # the new tip appends in place again
    assert k3.data_ptr() == k2.data_ptr() and torch.equal(k3, torch.cat([k2, c1[:, :, :1]], 2))


def test_two_appends_to_the_same_base():
    """torch.cat semantics: two different chunks appended to the SAME base give two independent results, whatever the
    base is (a plain tensor, or the latest / an older result of kv_append)."""
    B, H, D = 2, 3, 64
    base_k, base_v = _rand((B, H, 6, D), "bf16", 1), _rand((B, H, 6, D), "bf16", 2)
    a_k, a_v = _rand((B, H, 2, D), "bf16", 3), _rand((B, H, 2, D), "bf16", 4)
    b_k, b_v = _rand((B, H, 3, D), "bf16", 5), _rand((B, H, 3, D), "bf16", 6)
    for bk, bv in ((base_k, base_v), kernels.kv_append(base_k, base_v, a_k, a_v)):
        want_k, want_v = bk.clone(), bv.clone()
        r1 = kernels.kv_append(bk, bv, a_k, a_v)
        r2 = kernels.kv_append(bk, bv, b_k, b_v)
        assert torch.equal(r1[0], torch.cat([want_k, a_k], 2)) and torch.equal(r1[1], torch.cat([want_v, a_v], 2))
        assert torch.equal(r2[0], torch.cat([want_k, b_k], 2)) and torch.equal(r2[1], torch.cat([want_v, b_v], 2))
        assert torch.equal(bk, want_k) and torch.equal(bv, want_v)
        # and the masked form
        mask = torch.tensor([[1, 0, 1], [1, 1, 0]], dtype=torch.uint8, device="cuda")
        alen = torch.tensor([2, 1], dtype=torch.int32, device="cuda")
        m1 = kernels.registry.get_best("kv_append_with_mask", "cuda")(bk, bv, b_k, b_v, mask, alen)
        assert torch.equal(r1[0], torch.cat([want_k, a_k], 2)), "an earlier result changed under a later append"
        L0 = bk.shape[2]
        assert torch.equal(m1[0][0, :, L0], b_k[0, :, 0]) and torch.equal(m1[0][0, :, L0 + 1], b_k[0, :, 2])
        assert torch.equal(m1[0][:, :, :L0], want_k)


def test_misaligned_base_pointers_use_narrower_vectors():
    """views whose strides are 16-byte multiples but whose storage offset is not (a head-dim slice of a fused
    QKV buffer) must not reach the 16-byte kernels -- a misaligned 128-bit access is a sticky context error."""
    B, H, L, R, D = 2, 3, 5, 4, 64
    fused_new = _rand((B, H, R, 3 * D + 8), "bf16", 1)       # row stride 400 B (16-byte multiple)
    fused_base = _rand((B, H, L, 3 * D + 8), "bf16", 2)
    for off in (4, 1, 2):  # 8-, 2-, 4-byte storage offsets
        nk, nv = fused_new[..., off:off + D], fused_new[..., off + D:off + 2 * D]
        bk, bv = fused_base[..., off:off + D], fused_base[..., off + D:off + 2 * D]
        ok, ov = kernels.kv_append(bk, bv, nk, nv)
        assert torch.equal(ok, torch.cat([bk, nk], 2)) and torch.equal(ov, torch.cat([bv, nv], 2)), off
    torch.cuda.synchronize()


def test_strided_and_odd_layouts():
    # new rows given as a [:, :, :n] slice of a wider tensor (slice_prefix view), head dim 20 (8-byte rows in bf16 -> 8B vectors)
    for D, dtype in [(20, "bf16"), (6, "fp16"), (3, "fp32"), (1, "bf16"), (128, "bf16")]:
        B, H, L = 2, 3, 7
        bk, bv = _rand((B, H, L, D), dtype, 1), _rand((B, H, L, D), dtype, 2)
        wide_k, wide_v = _rand((B, H, 6, D), dtype, 3), _rand((B, H, 6, D), dtype, 4)
        ok, ov = kernels.kv_append(bk, bv, wide_k[:, :, :4], wide_v[:, :, :4])
        assert torch.equal(ok, torch.cat([bk, wide_k[:, :, :4]], 2)), (D, dtype)
        assert torch.equal(ov, torch.cat([bv, wide_v[:, :, :4]], 2)), (D, dtype)
    # transposed base (non-contiguous head dim) goes through the layout fix
    bk = _rand((2, 3, 16, 7), "fp32", 9).transpose(2, 3)
    nk = _rand((2, 3, 2, 16), "fp32", 10)
    ok, _ = kernels.kv_append(bk, bk, nk, nk)
    assert torch.equal(ok, torch.cat([bk, nk], 2))


def test_empty_cases():
    bk = torch.zeros(2, 2, 0, 8, device="cuda")
    nk = torch.randn(2, 2, 3, 8, device="cuda")
    ok, ov = kernels.kv_append(bk, bk, nk, nk)
    assert torch.equal(ok, nk)
    ok2, _ = kernels.kv_append(ok, ov, nk[:, :, :0], nk[:, :, :0])
    assert torch.equal(ok2, nk)
    e = torch.zeros(0, 2, 4, 8, device="cuda")
    ok3, _ = kernels.kv_append(e, e, e[:, :, :2], e[:, :, :2])
    assert ok3.shape == (0, 2, 6, 8)


@pytest.mark.parametrize("variant", [1, 2, 3, (3, "early"), 3 + 16 * 2, (3 + 16 * 2, "early")],
                         ids=["vec", "bulk", "column", "column_early_rows", "column_rb2", "column_rb2_early_rows"])
@pytest.mark.parametrize("zero_rejected", [True, False])
def test_arena_all_layers_vs_oracle(variant, zero_rejected):
    """sd_kv_append over every layer with ragged lengths and device-side accept lengths."""
    nl, B, H, D, C, R = 3, 5, 4, 64, 48, 5
    arena = KVArena(nl, B, H, D, C, R, torch.bfloat16, "cuda")
    g = torch.Generator().manual_seed(4)
    arena.k.copy_(torch.randn(arena.k.shape, generator=g).to(torch.bfloat16))
    arena.v.copy_(torch.randn(arena.v.shape, generator=g).to(torch.bfloat16))
    lens = torch.tensor([0, 7, 33, 20, 1], dtype=torch.int32)
    arena.set_lengths(lens)
    exp_k, exp_v = to_np(arena.k).copy(), to_np(arena.v).copy()
    early = isinstance(variant, tuple)
    variant = variant[0] if early else variant
    try:
        cuda_ops.set_kv_variant(variant)
        cur = lens.numpy().copy()
        for step in range(3):
            arena.new_k.copy_(torch.randn(arena.new_k.shape, generator=g).to(torch.bfloat16))
            arena.new_v.copy_(torch.randn(arena.new_v.shape, generator=g).to(torch.bfloat16))
            accept = torch.randint(0, R, (B,), generator=g).to(torch.int32)
            if step == 0:
                accept[0], accept[1] = 0, R - 1
            fits = cur + R <= C
            accept_dev = accept.cuda()
            if not fits.all():
                break
            if early:  # contract of rows_ready_early: the staged rows were not written by the previous kernel
                torch.cuda.synchronize()
            arena.append(accept_dev, keep_bias=1, zero_rejected=zero_rejected, rows_ready_early=early)
            nk, nv = to_np(arena.new_k), to_np(arena.new_v)
            for l in range(nl):
                new_lens = c_oracle.kv_arena_append(exp_k[l], cur, nk[l], accept.numpy() + 1, zero_rejected)
                c_oracle.kv_arena_append(exp_v[l], cur, nv[l], accept.numpy() + 1, zero_rejected)
            cur = new_lens
            torch.cuda.synchronize()
            np.testing.assert_array_equal(arena.seq_lens.cpu().numpy(), cur)
            np.testing.assert_array_equal(to_np(arena.k), exp_k)
            np.testing.assert_array_equal(to_np(arena.v), exp_v)
        assert arena.overflow.item() == 0
    finally:
        cuda_ops.set_kv_variant(0)


def test_arena_equals_masked_reference_semantics(golden_dir):
    """arena[:, :, :L+K] after an append with prefix acceptance == kv_append_with_mask_ref output."""
    z = np.load(golden_dir / "kv_append.npz")
    name = "cat_bf16_2x2x33x64_k4"
    bk, bv, nk, nv = (from_np(z[f"{name}.{k}"], "bf16").cuda() for k in ("bk", "bv", "nk", "nv"))
    B, H, L, D = bk.shape
    K = nk.shape[2]
    arena = KVArena(1, B, H, D, L + K + 5, K, torch.bfloat16, "cuda")
    arena.k[0, :, :, :L] = bk
    arena.v[0, :, :, :L] = bv
    arena.set_lengths([L] * B)
    arena.new_k[0].copy_(nk)
    arena.new_v[0].copy_(nv)
    arena.append(torch.from_numpy(z[f"{name}.alen"]).cuda(), keep_bias=0, zero_rejected=True)
    np.testing.assert_array_equal(to_np(arena.k[0][:, :, : L + K]), z[f"{name}.mk"])
    np.testing.assert_array_equal(to_np(arena.v[0][:, :, : L + K]), z[f"{name}.mv"])
    np.testing.assert_array_equal(arena.seq_lens.cpu().numpy(), L + z[f"{name}.alen"])


def test_capacity_is_enforced():
    arena = KVArena(1, 2, 2, 16, 8, 5, torch.bfloat16, "cuda")
    arena.set_lengths([4, 4])
    with pytest.raises(RuntimeError, match="capacity"):
        arena.append(None)
    # device-side guard when the host bound is unknown: rows beyond capacity are dropped and flagged
    arena.set_lengths([6, 1], max_len_host=0)
    arena.new_k.fill_(1.0)
    arena.new_v.fill_(1.0)
    arena.append(None, check_capacity=False)
    torch.cuda.synchronize()
    assert arena.overflow.item() == 1
    assert arena.seq_lens.tolist() == [8, 6]


def test_llama_shape_round_trip():
    """Full BASELINE shape (28 layers, B=64, Hkv=8, D=128, R=5): append then read back through the
    per-sequence lengths -- the appended rows equal the staged rows, everything else is untouched."""
    nl, B, H, D, C, R = 28, 64, 8, 128, 64, 5
    arena = KVArena(nl, B, H, D, C, R, torch.bfloat16, "cuda")
    g = torch.Generator(device="cuda").manual_seed(1234)
    arena.new_k.copy_(torch.randn(arena.new_k.shape, generator=g, device="cuda").to(torch.bfloat16))
    arena.new_v.copy_(torch.randn(arena.new_v.shape, generator=g, device="cuda").to(torch.bfloat16))
    lens = torch.randint(0, C - R, (B,), dtype=torch.int32)
    arena.set_lengths(lens)
    accept = torch.randint(0, R, (B,), dtype=torch.int32).cuda()
    arena.append(accept, keep_bias=1)
    torch.cuda.synchronize()
    keep = (accept + 1).cpu()
    assert torch.equal(arena.seq_lens.cpu(), lens + keep)
    rows = torch.arange(C)[None, :]
    valid = (rows >= lens[:, None]) & (rows < (lens + keep)[:, None])          # [B, C]
    zeroed = (rows >= (lens + keep)[:, None]) & (rows < (lens + R)[:, None])
    k = arena.k.cpu()
    for b in range(0, B, 7):
        got = k[:, b][:, :, valid[b]]
        want = arena.new_k[:, b, :, : int(keep[b])].cpu()
        assert torch.equal(got, want)
        assert not k[:, b][:, :, zeroed[b]].any()
    kept_nonzero = sum(int(arena.new_k[:, b, :, : int(keep[b])].ne(0).sum()) for b in range(B))
    assert int(k.ne(0).sum()) == kept_nonzero  # nothing outside the appended rows was written
