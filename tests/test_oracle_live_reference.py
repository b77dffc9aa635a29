"""CPU, build container only: the oracle against the UNMODIFIED reference module, live, on randomised inputs.

tests/test_oracle.py pins the oracle to committed fixtures (tests/golden/*.npz, generated once by running the
reference).  This file widens that pin where the reference checkout is present: the reference's own
src/kernels/reference.py is loaded from the file where it lies (by path, under a private module name -- this repo's
drop-in package is also called ``kernels``) and compared with all three oracle restatements on a few hundred seeded
cases per function: ragged shapes, ties, NaN / +-inf / -0.0 rows, every logits dtype, prefix and non-prefix masks,
accept lengths beyond the mask.  /root/reference does not exist on the GPU box: the whole file is skipped there.
Bit-exact everywhere (integer / index / byte work).
"""
import importlib.util
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import specdec_oracle as O
from oracle import torch_port

REF_FILE = Path(os.environ.get("SD_REFERENCE_ROOT", "/root/reference")) / "src" / "kernels" / "reference.py"
pytestmark = pytest.mark.skipif(not REF_FILE.exists(), reason="reference checkout not present (GPU box)")

TORCH_DT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("_live_reference_kernels", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16).copy()
    return t.numpy().copy()


def _logits_case(rng: np.random.Generator, gen: torch.Generator, dtype: str):
    B = int(rng.integers(1, 7))
    K = int(rng.integers(1, 9))
    V = int(rng.choice([2, 3, 17, 64, 257, 1000, 4099]))
    x = torch.randn(B, K, V, generator=gen) * float(rng.choice([0.05, 1.0, 30.0]))
    ids = torch.randint(0, V, (B, K), generator=gen)
    kind = int(rng.integers(0, 6))
    if kind >= 1:  # plant acceptance so that prefixes of every length occur
        plant = torch.rand(B, K, generator=gen) < 0.75
        bi, ki = torch.nonzero(plant, as_tuple=True)
        x[bi, ki, ids[bi, ki]] = x.abs().max() + 1.0
    if kind == 2:  # exact ties: the draft token ties with an EARLIER and a LATER index
        for b in range(B):
            for k in range(K):
                t = int(ids[b, k])
                other = int(rng.integers(0, V))
                x[b, k, other] = x[b, k, t]
    if kind == 3:  # quantisation ties: many equal values after the cast
        x = (x * 4).round() / 4
    if kind == 4:  # specials
        for _ in range(int(rng.integers(1, 6))):
            b, k, v = int(rng.integers(0, B)), int(rng.integers(0, K)), int(rng.integers(0, V))
            x[b, k, v] = float(rng.choice([np.nan, np.inf, -np.inf, -0.0, 0.0]))
    if kind == 5:  # a constant row (argmax = first index) and an all -inf row
        x[0, 0, :] = 0.0
        x[-1, -1, :] = -np.inf
    return x.to(TORCH_DT[dtype]), ids


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_verify_prefix_live(ref, dtype):
    rng = np.random.default_rng(20260 + len(dtype))
    gen = torch.Generator().manual_seed(1234)
    lens_seen = set()
    for case in range(150):
        logits, ids = _logits_case(rng, gen, dtype)
        want_len, want_mask = ref.verify_prefix_ref(logits, ids)
        want_len, want_mask = want_len.numpy(), want_mask.numpy()
        lens_seen.update(int(v) for v in want_len)
        raw = _bits(logits)
        got = O.verify_prefix(raw, ids.numpy(), dtype)
        np.testing.assert_array_equal(got[0], want_len, err_msg=f"numpy oracle, case {case}")
        np.testing.assert_array_equal(got[1], want_mask, err_msg=f"numpy oracle, case {case}")
        c_len, c_mask, c_tgt, _ = c_oracle.verify(raw, dtype, ids.numpy())
        np.testing.assert_array_equal(c_len, want_len, err_msg=f"C oracle, case {case}")
        np.testing.assert_array_equal(c_mask, want_mask, err_msg=f"C oracle, case {case}")
        np.testing.assert_array_equal(c_tgt, torch.argmax(logits, dim=-1).numpy(), err_msg=f"C oracle argmax, case {case}")
        p_len, p_mask = torch_port.verify_prefix_port(logits, ids)
        np.testing.assert_array_equal(p_len.numpy(), want_len, err_msg=f"torch port, case {case}")
        np.testing.assert_array_equal(p_mask.numpy(), want_mask, err_msg=f"torch port, case {case}")
        assert got[0].dtype == np.int32 and got[1].dtype == np.uint8
        assert want_len.dtype == np.int32 and want_mask.dtype == np.uint8
    assert {0, 1, 2}.issubset(lens_seen) and max(lens_seen) >= 6, "the cases must exercise short and long prefixes"


def _kv_case(rng: np.random.Generator, gen: torch.Generator, dtype: str):
    B, H = int(rng.integers(1, 5)), int(rng.integers(1, 4))
    L, K, D = int(rng.integers(0, 12)), int(rng.integers(1, 9)), int(rng.choice([1, 8, 64, 80]))
    dt = TORCH_DT[dtype]
    mk = lambda *s: torch.randn(*s, generator=gen).to(dt)
    return mk(B, H, L, D), mk(B, H, L, D), mk(B, H, K, D), mk(B, H, K, D)


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_kv_append_live(ref, dtype):
    rng = np.random.default_rng(7 + len(dtype))
    gen = torch.Generator().manual_seed(99)
    for case in range(60):
        bk, bv, nk, nv = _kv_case(rng, gen, dtype)
        want_k, want_v = ref.kv_append_ref(bk, bv, nk, nv)
        got_k, got_v = O.kv_append(_bits(bk), _bits(bv), _bits(nk), _bits(nv))
        np.testing.assert_array_equal(got_k, _bits(want_k), err_msg=f"case {case}")
        np.testing.assert_array_equal(got_v, _bits(want_v), err_msg=f"case {case}")
        np.testing.assert_array_equal(c_oracle.kv_append(_bits(bk), _bits(nk)), _bits(want_k), err_msg=f"C oracle, case {case}")
        p_k, p_v = torch_port.kv_append_port(bk, bv, nk, nv)
        assert torch.equal(p_k, want_k) and torch.equal(p_v, want_v)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_kv_append_with_mask_live(ref, dtype):
    """Masked, compacting append (reference.py:96-159): prefix masks (what verify produces), arbitrary masks, accept
    lengths smaller / larger than the number of set bits, zero acceptance; and the arena form on prefix masks."""
    rng = np.random.default_rng(31 + len(dtype))
    gen = torch.Generator().manual_seed(5)
    for case in range(120):
        bk, bv, dk, dv = _kv_case(rng, gen, dtype)
        B, H, L, D = bk.shape
        K = dk.shape[2]
        prefix = case % 2 == 0
        if prefix:
            alen = torch.from_numpy(rng.integers(0, K + 1, size=B).astype(np.int32))
            mask = (torch.arange(K)[None, :] < alen[:, None]).to(torch.uint8)
        else:
            mask = torch.from_numpy((rng.random((B, K)) < 0.6).astype(np.uint8))
            alen = torch.from_numpy(rng.integers(0, K + 2, size=B).astype(np.int32))
        want_k, want_v = ref.kv_append_with_mask_ref(bk, bv, dk, dv, mask, alen)
        got_k, got_v = O.kv_append_with_mask(_bits(bk), _bits(bv), _bits(dk), _bits(dv), mask.numpy(), alen.numpy())
        np.testing.assert_array_equal(got_k, _bits(want_k), err_msg=f"numpy oracle, case {case}")
        np.testing.assert_array_equal(got_v, _bits(want_v), err_msg=f"numpy oracle, case {case}")
        np.testing.assert_array_equal(c_oracle.kv_append(_bits(bk), _bits(dk), mask.numpy(), alen.numpy()), _bits(want_k),
                                      err_msg=f"C oracle, case {case}")
        if prefix:  # the in-place arena layout of the B200 path reproduces the same bytes
            C_ = L + K + 3
            ak = np.full((B, H, C_, D), 0x7F, dtype=_bits(bk).dtype)
            av = ak.copy()
            ak[:, :, :L], av[:, :, :L] = _bits(bk), _bits(bv)
            new_len = O.kv_arena_append(ak, av, np.full(B, L), _bits(dk), _bits(dv), alen.numpy(), zero_rejected=True)
            np.testing.assert_array_equal(ak[:, :, : L + K], _bits(want_k), err_msg=f"arena, case {case}")
            np.testing.assert_array_equal(av[:, :, : L + K], _bits(want_v), err_msg=f"arena, case {case}")
            np.testing.assert_array_equal(new_len, L + alen.numpy())
            assert (ak[:, :, L + K :] == 0x7F).all(), "the arena form must not touch rows beyond the offered slots"


def test_reference_shape_errors_are_assertion_errors(ref):
    """The drop-in raises what the reference raises (AssertionError) for the reference's own checks."""
    x = torch.zeros(2, 3, 5)
    with pytest.raises(AssertionError):
        ref.verify_prefix_ref(x, torch.zeros(2, 4, dtype=torch.long))
    with pytest.raises(AssertionError):
        O.verify_prefix(x.numpy(), np.zeros((2, 4), np.int64), "fp32")
    with pytest.raises(AssertionError):
        ref.verify_prefix_ref(x[0], torch.zeros(2, 3, dtype=torch.long))
    with pytest.raises(AssertionError):
        O.verify_prefix(x[0].numpy(), np.zeros((2, 3), np.int64), "fp32")


def test_oracle_pinned_to_vectors_regenerated_with_a_fresh_seed(tmp_path):
    """Everything tests/golden/make_golden.py records -- verify, kv_append, realign, the four policies, the bonus
    sampler's distribution -- regenerated from the live reference with ANOTHER seed into a scratch directory, and the
    whole of tests/test_oracle.py run against those vectors: the oracle is pinned to the reference, not to one set of
    committed numbers.  (Subprocesses: the reference's packages share their names with this repo's drop-ins.)"""
    import subprocess
    import sys

    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, SD_GOLDEN_OUT=str(tmp_path), SD_GOLDEN_SEED="4321")
    gen = subprocess.run([sys.executable, str(root / "tests" / "golden" / "make_golden.py")], env=env, capture_output=True,
                         text=True, timeout=600)
    assert gen.returncode == 0, gen.stderr[-2000:]
    assert {p.name for p in tmp_path.glob("*.npz")} >= {"verify.npz", "kv_append.npz", "kv_realign.npz", "policies.npz", "sampler.npz"}
    env = dict(os.environ, SD_GOLDEN_DIR=str(tmp_path))
    run = subprocess.run([sys.executable, "-m", "pytest", str(root / "tests" / "test_oracle.py"), "-q", "-x", "-p", "no:cacheprovider"],
                         env=env, capture_output=True, text=True, timeout=600, cwd=str(root))
    assert run.returncode == 0, run.stdout[-3000:]
    assert " passed" in run.stdout and "failed" not in run.stdout
