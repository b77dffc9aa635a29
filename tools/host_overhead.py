#!/usr/bin/env python3
"""Host cost per call of the registry drop-ins (the reference calls verify_prefix once per sequence,
src/specdec/core/pipeline.py:3018, and kv_append once per layer, src/specdec/models/hf_wrappers.py:1009).

    python tools/host_overhead.py

N calls are issued back to back without synchronising (tiny inputs: the GPU keeps up), then one synchronise: the
quotient is the host time per call (Python + ctypes + launch).  Printed next to the same measurement for the incumbents
(torch.argmax / torch.cat) and to the wall time per call WITH a synchronise on both sides (microbench protocol)."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import torch  # noqa: E402

import kernels  # noqa: E402


def per_call(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return {"host_us_per_call": round((t1 - t0) / n * 1e6, 2), "with_final_sync_us_per_call": round((t2 - t0) / n * 1e6, 2)}


def synced(fn, n=200):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return round(ts[len(ts) // 2] * 1e6, 2)


def main():
    out = {"device": torch.cuda.get_device_name(0)}
    logits = torch.randn(1, 4, 50257, device="cuda").to(torch.bfloat16)
    ids = torch.randint(0, 50257, (1, 4), device="cuda")
    out["verify_prefix B=1 K=4 V=50257"] = dict(per_call(lambda: kernels.verify_prefix(logits, ids)), synced_us=synced(lambda: kernels.verify_prefix(logits, ids)))
    out["torch.argmax (incumbent, no host loop)"] = dict(per_call(lambda: torch.argmax(logits, -1)), synced_us=synced(lambda: torch.argmax(logits, -1)))
    B, H, L, D = 1, 8, 1024, 128
    k = torch.randn(B, H, L, D, device="cuda").to(torch.bfloat16)
    nk = torch.randn(B, H, 4, D, device="cuda").to(torch.bfloat16)
    state = {"k": kernels.kv_append(k, k, nk, nk)}

    def ours():
        ck, cv = state["k"]
        if ck.shape[2] > 60000:  # restart before the arena grows without bound
            state["k"] = kernels.kv_append(k, k, nk, nk)
            ck, cv = state["k"]
        state["k"] = kernels.kv_append(ck, cv, nk, nk)

    out["kv_append B=1 H=8 L=1024 D=128 Kn=4 (cache = kv_append(cache, chunk))"] = dict(per_call(ours), synced_us=synced(ours))
    out["torch.cat (incumbent, same shapes, O(L))"] = dict(per_call(lambda: (torch.cat([k, nk], 2), torch.cat([k, nk], 2))),
                                                            synced_us=synced(lambda: (torch.cat([k, nk], 2), torch.cat([k, nk], 2))))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
