"""Shared test helpers: torch <-> numpy (bf16 as uint16 bits) and seeded generators of the
reference's synthetic inputs (scripts/microbench_verify.py:57-64 planted-match protocol)."""
import numpy as np
import torch

TORCH_DT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
SEED = 1234  # reference default seed (src/specdec/core/pipeline.py:407)


def to_np(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def from_np(a: np.ndarray, dtype: str) -> torch.Tensor:
    if dtype == "bf16":
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int16).copy()).view(torch.bfloat16)
    return torch.from_numpy(np.ascontiguousarray(a).copy())


def np_dtype_name(t: torch.Tensor) -> str:
    return {torch.float32: "fp32", torch.float16: "fp16", torch.bfloat16: "bf16"}[t.dtype]


def make_verify_inputs(B, rows, K, V, dtype="bf16", p_match=0.5, seed=SEED, device="cpu", id_dtype=torch.int64):
    """randn logits with planted 10.0 at the draft token with probability p_match."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, rows, V, generator=g)
    ids = torch.randint(0, V, (B, K), generator=g, dtype=torch.int64)
    plant = torch.rand(B, K, generator=g) < p_match
    bi, ki = torch.nonzero(plant, as_tuple=True)
    logits[bi, ki, ids[bi, ki]] = 10.0
    return logits.to(TORCH_DT[dtype]).to(device), ids.to(id_dtype).to(device)
