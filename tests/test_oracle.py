"""Pins oracle/ (numpy, C and the torch port) against the golden vectors produced by the
reference's own Python implementation (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import specdec_oracle as O
from oracle import torch_port as P

NP_TO_TORCH = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


def to_torch(a: np.ndarray, dtype: str) -> torch.Tensor:
    if dtype == "bf16":
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)
    return torch.from_numpy(a.copy())


@pytest.fixture(scope="module")
def verify_golden(golden_dir):
    return np.load(golden_dir / "verify.npz")


@pytest.fixture(scope="module")
def kv_golden(golden_dir):
    return np.load(golden_dir / "kv_append.npz")


def _verify_cases(z):
    for entry in z["__cases__"]:
        name, dtype = str(entry).split("|")
        yield name, dtype


def test_verify_numpy_oracle_matches_reference(verify_golden):
    z = verify_golden
    n = 0
    for name, dtype in _verify_cases(z):
        acc, mask = O.verify_prefix(z[f"{name}.logits"], z[f"{name}.ids"], dtype)
        assert acc.dtype == np.int32 and mask.dtype == np.uint8
        np.testing.assert_array_equal(acc, z[f"{name}.accept_len"], err_msg=name)
        np.testing.assert_array_equal(mask, z[f"{name}.mask"], err_msg=name)
        np.testing.assert_array_equal(O.argmax_first(O.as_f32(z[f"{name}.logits"], dtype)), z[f"{name}.argmax"], err_msg=name)
        n += 1
    assert n >= 25


def test_verify_c_oracle_matches_reference(verify_golden):
    z = verify_golden
    for name, dtype in _verify_cases(z):
        acc, mask, tgt, _ = c_oracle.verify(z[f"{name}.logits"], dtype, z[f"{name}.ids"])
        np.testing.assert_array_equal(acc, z[f"{name}.accept_len"], err_msg=name)
        np.testing.assert_array_equal(mask, z[f"{name}.mask"], err_msg=name)
        np.testing.assert_array_equal(tgt, z[f"{name}.argmax"], err_msg=name)


def test_verify_torch_port_matches_reference(verify_golden):
    z = verify_golden
    for name, dtype in _verify_cases(z):
        acc, mask = P.verify_prefix_port(to_torch(z[f"{name}.logits"], dtype), torch.from_numpy(z[f"{name}.ids"]))
        np.testing.assert_array_equal(acc.numpy(), z[f"{name}.accept_len"], err_msg=name)
        np.testing.assert_array_equal(mask.numpy(), z[f"{name}.mask"], err_msg=name)


def test_adversarial_expectations(verify_golden):
    """The semantics we rely on, read off the reference's outputs: first max on ties, NaN maximal
    (first NaN), -0.0 == +0.0."""
    z = verify_golden
    for dtype in ("bf16", "fp32", "fp16"):
        am = z[f"adversarial_{dtype}.argmax"].reshape(-1)
        assert am[0] == 17 and am[1] == 0 and am[2] == 77 and am[3] == 4000
        assert am[4] == 100 and am[5] == 50 and am[6] == 0 and am[7] == 4098 and am[8] == 0


def test_verify_bonus_consistency(verify_golden):
    """K+1-row form: rows 0..K-1 behave like the reference; bonus = argmax of row accept_len."""
    z = verify_golden
    for name, dtype in _verify_cases(z):
        lg, ids = z[f"{name}.logits"], z[f"{name}.ids"]
        if ids.shape[1] < 2:
            continue
        acc, mask, tgt, bonus = O.verify_prefix_bonus(lg, ids[:, :-1], dtype)
        ref_acc = np.minimum(z[f"{name}.accept_len"], ids.shape[1] - 1)
        np.testing.assert_array_equal(acc, ref_acc)
        np.testing.assert_array_equal(tgt, z[f"{name}.argmax"])
        np.testing.assert_array_equal(bonus, z[f"{name}.argmax"][np.arange(len(acc)), acc])
        c_acc, c_mask, c_tgt, c_bonus = c_oracle.verify(lg, dtype, ids[:, :-1])
        np.testing.assert_array_equal(c_acc, acc)
        np.testing.assert_array_equal(c_mask, mask)
        np.testing.assert_array_equal(c_bonus, bonus)
        tp_acc, _, tp_bonus = P.verify_step_port(to_torch(lg, dtype), torch.from_numpy(ids[:, :-1]))
        np.testing.assert_array_equal(tp_acc.numpy(), acc)
        np.testing.assert_array_equal(tp_bonus.numpy(), bonus)


def test_verify_shape_errors():
    with pytest.raises(AssertionError):
        O.verify_prefix(np.zeros((2, 3, 8), np.float32), np.zeros((2, 2), np.int64))
    with pytest.raises(AssertionError):
        O.verify_prefix(np.zeros((2, 3, 8), np.float32), np.zeros((3, 3), np.int64))
    with pytest.raises(AssertionError):
        O.verify_prefix(np.zeros((3, 8), np.float32), np.zeros((3,), np.int64))


def _kv_cases(z):
    for entry in z["__cases__"]:
        name, dtype, _ = str(entry).split("|")
        yield name, dtype


def test_kv_append_oracles_match_reference(kv_golden):
    z = kv_golden
    n = 0
    for name, dtype in _kv_cases(z):
        bk, bv, nk, nv = (z[f"{name}.{k}"] for k in ("bk", "bv", "nk", "nv"))
        ok, ov = O.kv_append(bk, bv, nk, nv)
        np.testing.assert_array_equal(ok, z[f"{name}.ok"])
        np.testing.assert_array_equal(ov, z[f"{name}.ov"])
        np.testing.assert_array_equal(c_oracle.kv_append(bk, nk), z[f"{name}.ok"])
        tk, tv = P.kv_append_port(to_torch(bk, dtype), to_torch(bv, dtype), to_torch(nk, dtype), to_torch(nv, dtype))
        assert tuple(tk.shape) == z[f"{name}.ok"].shape
        # masked
        mk, mv = O.kv_append_with_mask(bk, bv, nk, nv, z[f"{name}.mask"], z[f"{name}.alen"])
        np.testing.assert_array_equal(mk, z[f"{name}.mk"])
        np.testing.assert_array_equal(mv, z[f"{name}.mv"])
        np.testing.assert_array_equal(c_oracle.kv_append(bv, nv, z[f"{name}.mask"], z[f"{name}.alen"]), z[f"{name}.mv"])
        if f"{name}.hk" in z:
            hk, hv = O.kv_append_with_mask(bk, bv, nk, nv, z[f"{name}.hole_mask"], z[f"{name}.hole_alen"])
            np.testing.assert_array_equal(hk, z[f"{name}.hk"])
            np.testing.assert_array_equal(hv, z[f"{name}.hv"])
            np.testing.assert_array_equal(c_oracle.kv_append(bk, nk, z[f"{name}.hole_mask"], z[f"{name}.hole_alen"]), z[f"{name}.hk"])
        n += 1
    assert n == 18


def test_kv_arena_append_equals_masked_reference(kv_golden):
    """The in-place arena form reproduces kv_append_with_mask_ref for prefix masks."""
    z = kv_golden
    for name, dtype in _kv_cases(z):
        bk, nk = z[f"{name}.bk"], z[f"{name}.nk"]
        B, H, L, D = bk.shape
        K = nk.shape[2]
        cap = L + K + 3
        for impl in ("numpy", "c"):
            arena_k = np.full((B, H, cap, D), 7, dtype=bk.dtype)
            arena_v = np.full((B, H, cap, D), 7, dtype=bk.dtype)
            arena_k[:, :, :L] = bk
            arena_v[:, :, :L] = z[f"{name}.bv"]
            lens = np.full(B, L, np.int32)
            if impl == "numpy":
                new_lens = O.kv_arena_append(arena_k, arena_v, lens, nk, z[f"{name}.nv"], z[f"{name}.alen"])
            else:
                new_lens = c_oracle.kv_arena_append(arena_k, lens, nk, z[f"{name}.alen"])
                c_oracle.kv_arena_append(arena_v, lens, z[f"{name}.nv"], z[f"{name}.alen"])
            np.testing.assert_array_equal(new_lens, L + z[f"{name}.alen"])
            np.testing.assert_array_equal(arena_k[:, :, : L + K], z[f"{name}.mk"])
            np.testing.assert_array_equal(arena_v[:, :, : L + K], z[f"{name}.mv"])
            assert (arena_k[:, :, L + K :] == 7).all()


def test_kv_append_shape_errors():
    a = np.zeros((2, 2, 3, 4), np.float32)
    for bad in (np.zeros((1, 2, 2, 4)), np.zeros((2, 4, 2, 4)), np.zeros((2, 2, 2, 8))):
        with pytest.raises(AssertionError):
            O.kv_append(a, a, bad.astype(np.float32), bad.astype(np.float32))


def test_kv_realign_matches_reference(golden_dir):
    z = np.load(golden_dir / "kv_realign.npz")
    for name in z["__cases__"]:
        name = str(name)
        rk, rv = O.kv_realign(z[f"{name}.k"], z[f"{name}.v"], z[f"{name}.new_lengths"])
        np.testing.assert_array_equal(rk, z[f"{name}.rk"])
        np.testing.assert_array_equal(rv, z[f"{name}.rv"])
        n0 = z[f"{name}.slice_k"].shape[2]
        np.testing.assert_array_equal(z[f"{name}.k"][:, :, :n0], z[f"{name}.slice_k"])


def test_policies_match_reference(golden_dir):
    z = np.load(golden_dir / "policies.npz")
    for name in z["__cases__"]:
        name = str(name)
        base, draft, ids = z[f"{name}.base"], z[f"{name}.draft"], z[f"{name}.ids"]
        K = ids.shape[1]
        for p, want in zip(z[f"{name}.typical_p"], z[f"{name}.typical"]):
            got, probs = O.accept_typical(base[:, :K], ids, float(p))
            assert got[0] == want, (name, p)
            ref_p = np.take_along_axis(z[f"{name}.base_probs"], ids[..., None], -1)[..., 0]
            np.testing.assert_allclose(probs, ref_p, rtol=1e-5, atol=1e-7)
        for t, want in zip(z[f"{name}.conf_tau"], z[f"{name}.conf"]):
            got, _ = O.accept_confidence(draft, float(t))
            assert got[0] == want, (name, t)
        V = base.shape[-1]
        for k, want in zip(z[f"{name}.topk_k"], z[f"{name}.topk"]):
            got = O.accept_topk(base[:, :K], ids, int(min(k, V)))
            assert got[0] == want, (name, k)
        acc, _, tgt, bonus = O.verify_prefix_bonus(base, ids)
        assert acc[0] == z[f"{name}.longest_prefix"][0]
        assert bonus[0] == z[f"{name}.bonus_greedy"][0]


def test_bonus_sampler_distribution_matches_reference(golden_dir):
    """do_sample=True: the distribution the reference hands to torch.multinomial (recorded by
    tests/golden/make_golden.py through a wrapper around torch.multinomial, pipeline.py:90-133) equals
    softmax(bonus_logits_filter(...)): same support, probabilities to fp32 rounding.  With exact ties across the
    nucleus boundary torch.sort's (unstable) order decides WHICH of the equal logits survive; the oracle and the
    kernel keep the lowest indices -- there the kept VALUES (sorted probabilities) and the support size must match."""
    z = np.load(golden_dir / "sampler.npz")
    n_exact = 0
    for name in z["__cases__"]:
        name = str(name)
        row = z[f"{name}.row"]
        T = float(z[f"{name}.T"][0])
        k = int(z[f"{name}.k"][0]) or None
        p = float(z[f"{name}.p"][0])
        p = None if p >= 1.0 else p
        ref = z[f"{name}.probs"]
        x = O.bonus_logits_filter(row, T, k, p)
        e = np.exp((x - x.max()).astype(np.float64))
        probs = e / e.sum()
        assert np.isfinite(x).sum() == (ref > 0).sum(), name
        assert ref[int(z[f"{name}.token"][0])] > 0, name  # the token the reference drew lies in its own support
        if np.array_equal(np.isfinite(x), ref > 0):
            np.testing.assert_allclose(probs, ref, rtol=2e-5, atol=1e-7, err_msg=name)
            n_exact += 1
        else:
            assert int(z[f"{name}.ties"][0]) == 1, f"{name}: support differs without ties in the row"
            np.testing.assert_allclose(np.sort(probs), np.sort(ref), rtol=2e-5, atol=1e-7, err_msg=name)
        # the inverse-CDF draw of the oracle lands inside the support for any u
        for u in (0.0, 0.37, 0.999999):
            tok, _ = O.sample_bonus(row, T, k, p, u)
            assert np.isfinite(x[tok]), (name, u)
    # committed fixtures (seed 1234): 26 of 28 rows have no tie across the nucleus boundary; vectors regenerated with
    # another seed (SD_GOLDEN_DIR, tests/test_oracle_live_reference.py) may have a few more tied rows
    import os

    assert n_exact >= (26 if "SD_GOLDEN_DIR" not in os.environ else len(z["__cases__"]) - 8)


def test_bf16_roundtrip_helper():
    x = torch.randn(4096, generator=torch.Generator().manual_seed(0))
    x[0], x[1], x[2] = float("nan"), float("inf"), -0.0
    want = x.bfloat16().view(torch.int16).numpy().view(np.uint16)
    got = O.f32_to_bf16_bits(x.numpy())
    ok = (got == want) | (np.isnan(x.numpy()))
    assert ok.all()
    np.testing.assert_array_equal(O.as_f32(want)[3:], x.bfloat16().float().numpy()[3:])


def test_attention_oracle_against_torch_sdpa():
    """parity unpinned by the reference (third-party SDPA); checked against torch's fp32 math SDPA."""
    g = torch.Generator().manual_seed(5)
    B, Hq, Hkv, R, D, cap = 2, 6, 2, 5, 16, 40
    lens = [9, 23]
    q = torch.randn(B, Hq, R, D, generator=g)
    k = torch.randn(B, Hkv, cap, D, generator=g)
    v = torch.randn(B, Hkv, cap, D, generator=g)
    got = O.verify_attention(q.numpy(), k.numpy(), v.numpy(), lens)
    for b in range(B):
        n = lens[b] + R
        kk = k[b : b + 1, :, :n].repeat_interleave(Hq // Hkv, dim=1)
        vv = v[b : b + 1, :, :n].repeat_interleave(Hq // Hkv, dim=1)
        mask = torch.arange(n)[None, :] <= (lens[b] + torch.arange(R))[:, None]
        want = torch.nn.functional.scaled_dot_product_attention(q[b : b + 1], kk, vv, attn_mask=mask)
        np.testing.assert_allclose(got[b], want[0].numpy(), rtol=1e-4, atol=1e-5)
