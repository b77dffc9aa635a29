"""CPU port of the reference's verify step using the same torch CPU arithmetic it uses.

TEST INFRASTRUCTURE ONLY (see oracle/specdec_oracle.py).  This is the arm that
``bench.py`` times as ``cpu_baseline`` / ``--impl reference`` ("kind": "port"):
the reference's own CPU path *is* torch-on-CPU (``torch.argmax`` + a Python
prefix loop, ``torch.cat`` per layer), and /root/reference cannot travel to the
GPU box, so the same torch calls are restated here and pinned against the
reference by tests/golden/ (tests/test_oracle.py).
"""

from __future__ import annotations

import torch


def verify_prefix_port(logits: torch.Tensor, draft_ids: torch.Tensor):
    """reference: src/kernels/reference.py:13-56.  One ``torch.argmax`` over V with
    all intra-op threads, then the per-sequence leading-run scan on the host."""
    assert logits.dim() == 3 and draft_ids.dim() == 2
    assert logits.shape[:2] == draft_ids.shape
    B, K, _ = logits.shape
    hit = (torch.argmax(logits, dim=-1) == draft_ids).tolist()
    accept_len = torch.zeros(B, dtype=torch.int32)
    mask = torch.zeros((B, K), dtype=torch.uint8)
    for b, row in enumerate(hit):
        n = 0
        while n < K and row[n]:
            n += 1
        accept_len[b] = n
        mask[b, :n] = 1
    return accept_len, mask


def verify_step_port(logits: torch.Tensor, draft_ids: torch.Tensor):
    """Verify over K+1 rows plus the greedy bonus token, the way the reference's
    batch loop obtains them: ``verify_prefix`` on the first K rows
    (src/specdec/policies/policies.py:128-134) and ``argmax`` of row accept_len
    (src/specdec/core/pipeline.py:3140-3163 / :135-137), one sequence at a time."""
    B, R, _ = logits.shape
    K = draft_ids.shape[1]
    accept_len, mask = verify_prefix_port(logits[:, :K], draft_ids)
    bonus = torch.empty(B, dtype=torch.int64)
    for b in range(B):
        row = min(int(accept_len[b]), R - 1)
        bonus[b] = torch.argmax(logits[b, row], dim=-1)
    return accept_len, mask, bonus


def kv_append_port(base_k, base_v, new_k, new_v):
    """reference: src/kernels/reference.py:59-93 -- out-of-place concatenation."""
    assert base_k.dim() == 4 and new_k.dim() == 4
    assert base_k.shape[0] == new_k.shape[0] and base_k.shape[1] == new_k.shape[1]
    assert base_k.shape[3] == new_k.shape[3]
    return torch.cat((base_k, new_k), dim=2), torch.cat((base_v, new_v), dim=2)


def kv_append_layers_port(base_layers, new_layers, accept_rows: int):
    """reference: src/specdec/models/hf_wrappers.py:985-1030 (``_append_kv_with_kernel``):
    per layer, append the accepted slice (``slice_prefix``: src/specdec/cache/kv_types.py:62-86)."""
    out = []
    for (bk, bv), (nk, nv) in zip(base_layers, new_layers):
        out.append(kv_append_port(bk, bv, nk[:, :, :accept_rows], nv[:, :, :accept_rows]))
    return out
