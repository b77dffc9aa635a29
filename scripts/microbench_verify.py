#!/usr/bin/env python3
"""Microbenchmark of the verify_prefix registry op -- the counterpart of the reference's
scripts/microbench_verify.py (same planted-match protocol, :57-64, its fp32 shapes :35-49) extended to the
shapes of BASELINE configs[1]: B = 1..256, K = 1..8, V = 50257, bf16.

    python scripts/microbench_verify.py [--quick]

Both arms run on the GPU: `kernels.verify_prefix` (this repo's registry entry) and the incumbent the
reference would use there (`torch.argmax` + its host prefix loop == verify_prefix_ref semantics, restated
inline so that nothing is imported from oracle/).  Times are wall-clock per call with a synchronise on both
sides, as in the reference script (so they include launch latency and the host loop -- that is the point)."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kernels import get_kernel_info, verify_prefix  # noqa: E402


def incumbent(logits, draft_ids):
    """What the reference runs on a CUDA tensor when no kernel is registered (reference.py:37-54)."""
    B, K, _ = logits.shape
    pred = torch.argmax(logits, dim=-1)
    acc = torch.zeros(B, dtype=torch.int32, device=logits.device)
    mask = torch.zeros((B, K), dtype=torch.uint8, device=logits.device)
    for b in range(B):
        n = 0
        for k in range(K):
            if pred[b, k].item() == draft_ids[b, k].item():
                n += 1
            else:
                break
        acc[b] = n
        mask[b, :n] = 1
    return acc, mask


def timeit(fn, n):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    info = get_kernel_info()
    print(json.dumps({"verify_backend": info["verify_backend"], "device": torch.cuda.get_device_name(0)}))
    rng = np.random.default_rng(1234)
    configs = [(1, 1, 4096, "fp32"), (1, 4, 4096, "fp32"), (8, 4, 4096, "fp32"), (1, 4, 32768, "fp32")]  # reference shapes
    bs = (1, 8, 64, 256) if a.quick else (1, 2, 4, 8, 16, 32, 64, 128, 256)
    ks = (1, 4, 8) if a.quick else (1, 2, 4, 8)
    configs += [(B, K, 50257, "bf16") for B in bs for K in ks]
    for B, K, V, dt in configs:
        dtype = torch.float32 if dt == "fp32" else torch.bfloat16
        logits = torch.randn(B, K, V)
        ids = torch.randint(0, V, (B, K))
        plant = torch.from_numpy(rng.random((B, K)) < 0.3)
        bi, ki = torch.nonzero(plant, as_tuple=True)
        logits[bi, ki, ids[bi, ki]] = 10.0
        logits, ids = logits.to(dtype).cuda(), ids.cuda()
        for _ in range(5):
            verify_prefix(logits, ids)
        got = verify_prefix(logits, ids)
        want = incumbent(logits, ids)
        ok = torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        t_k = timeit(lambda: verify_prefix(logits, ids), 50)
        t_r = timeit(lambda: incumbent(logits, ids), 5 if B * K > 256 else 20)
        print(json.dumps({"B": B, "K": K, "V": V, "dtype": dt, "kernel_us": round(t_k, 1), "torch_incumbent_us": round(t_r, 1),
                          "speedup": round(t_r / t_k, 1), "identical": ok, "logits_MB": round(B * K * V * logits.element_size() / 1e6, 2)}),
              flush=True)


if __name__ == "__main__":
    main()
