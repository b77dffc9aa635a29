"""GPU parity of the policy classes (drop-in for src/specdec/policies/policies.py) against the outputs of
the reference's own policy classes stored in tests/golden/policies.npz, and against the oracle on fresh
seeded inputs.  Accept lengths are exact unless a probability sits at the threshold."""
import numpy as np
import pytest
import torch

from oracle import specdec_oracle as O
from specdec.policies import create_policy
from specdec.policies.policies import TopKAgreementPolicy
from helpers import make_verify_inputs, to_np

pytestmark = pytest.mark.gpu


def test_reference_policy_outputs(golden_dir):
    z = np.load(golden_dir / "policies.npz")
    n = 0
    for name in z["__cases__"]:
        name = str(name)
        base = torch.from_numpy(z[f"{name}.base"]).cuda()
        draft = torch.from_numpy(z[f"{name}.draft"]).cuda()
        ids = torch.from_numpy(z[f"{name}.ids"]).cuda()
        base_tokens = torch.argmax(base, -1)
        probs = np.take_along_axis(z[f"{name}.base_probs"], z[f"{name}.ids"][..., None], -1)[..., 0]
        pmax = torch.softmax(draft.float(), -1).max(-1)[0].cpu().numpy()
        got, info = create_policy("longest_prefix").accept_tokens(ids, base_tokens, draft, base)
        assert got == int(z[f"{name}.longest_prefix"][0]) and info["policy"] == "longest_prefix" and info["accepted_len"] == got
        for p, want in zip(z[f"{name}.typical_p"], z[f"{name}.typical"]):
            got, info = create_policy("typical", p=float(p)).accept_tokens(ids, base_tokens, draft, base)
            if not (np.abs(probs - p) < 1e-4).any():
                assert got == want, (name, "typical", p)
            assert info["p"] == float(p) and info["proposed_len"] == ids.shape[1]
        for t, want in zip(z[f"{name}.conf_tau"], z[f"{name}.conf"]):
            got, info = create_policy("conf_threshold", tau=float(t)).accept_tokens(ids, base_tokens, draft, base)
            if not (np.abs(pmax - t) < 1e-4).any():
                assert got == want, (name, "conf", t)
        for k, want in zip(z[f"{name}.topk_k"], z[f"{name}.topk"]):
            got, info = create_policy("topk_agree", k=int(k)).accept_tokens(ids, base_tokens, draft, base)
            assert got == want, (name, "topk", k)
            assert info["k"] == int(k)
        n += 1
    assert n >= 3


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("k", [1, 5, 50])
def test_topk_agreement_batched_vs_oracle(dtype, k):
    B, K, V = 6, 4, 50257
    logits, ids = make_verify_inputs(B, K + 1, K, V, dtype, p_match=0.4, seed=31 + k, device="cuda")
    # make some draft tokens rank exactly k-1 / k (just inside / just outside the top-k)
    x = logits[:, :K].float()
    order = torch.argsort(x, dim=-1, descending=True)
    ids[0, 0], ids[1, 0] = order[0, 0, k - 1], order[1, 0, min(k, V - 1)]
    out = TopKAgreementPolicy(k).accept_batch(ids, base_logits=logits, proposed_logits=logits)
    want = O.accept_topk(to_np(logits), to_np(ids), k, dtype)
    if dtype == "fp32":  # no ties at the k-th place -> exact
        np.testing.assert_array_equal(out["accept_len"].cpu().numpy(), want)
    xs = O.as_f32(to_np(logits), dtype)[:, :K]
    greater = (xs > np.take_along_axis(xs, to_np(ids)[..., None], -1)).sum(-1)
    np.testing.assert_array_equal(out["token_prob"].cpu().numpy().astype(np.int64), greater)
    np.testing.assert_array_equal(out["accept_len"].cpu().numpy(), O._leading_run(greater < k))


def test_no_logits_falls_back_to_token_comparison():
    """reference: policies.py:181-196."""
    prop = torch.tensor([[5, 6, 7, 8]], device="cuda")
    base = torch.tensor([[5, 6, 9, 8]], device="cuda")
    got, info = create_policy("longest_prefix").accept_tokens(prop, base)
    assert got == 2 and info["accepted_len"] == 2


def test_unknown_policy_and_cpu_tensors_raise():
    with pytest.raises(ValueError):
        create_policy("nope")
    logits = torch.zeros(1, 2, 8)
    with pytest.raises(RuntimeError):
        create_policy("longest_prefix").accept_tokens(torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 2, dtype=torch.long), None, logits)
