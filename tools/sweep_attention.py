#!/usr/bin/env python3
"""Sweep tile configurations and split counts of the verify-attention kernel on one GPU.

    python tools/sweep_attention.py [--model llama3b|llama1b|gpt2|all] [--L 1024] [--iters 120]

One JSON line per (model, context, config, splits): microseconds per launch (CUDA events around graph replays
that rotate over enough layers for the K/V reads to come from HBM), achieved GB/s on the algorithmic bytes
(2*B*Hkv*(L+R)*D*2 KV + 2*B*Hq*R*D*2 Q/O) and the fraction of MEASURED_PEAKS.json.
config: 0 heuristic, 1 = 16-key tiles x 3 stages, 2 = 32-key x 3 (4 at D=64), 3 = 32-key x 2
(include/sd_verify.h, sd_verify_attention_set_tuning).  splits: 0 = the kernel's own heuristic.
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import torch  # noqa: E402

from kernels import cuda_ops  # noqa: E402

MODELS = {  # layers, Hq, Hkv, D
    "llama3b": (28, 24, 8, 128),
    "llama1b": (16, 32, 8, 64),
    "gpt2": (12, 12, 12, 64),
}


def peak_gbs() -> float:
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        return 6444.0  # the value measured on this pool in round 1


def time_replays(fns, iters):
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for fn in fns:
                fn()
    torch.cuda.synchronize()
    reps = max(2, iters // len(fns))
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * len(fns)) * 1e3


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="all")
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--L", type=int, nargs="*", default=[1024])
    ap.add_argument("--iters", type=int, default=120)
    ap.add_argument("--configs", type=int, nargs="*", default=[0, 1, 2, 3])
    ap.add_argument("--splits", type=int, nargs="*", default=[0, 1, 2, 3, 4, 6, 8])
    a = ap.parse_args()
    dev = torch.device("cuda")
    peak = peak_gbs()
    names = list(MODELS) if a.model == "all" else [a.model]
    R = a.K + 1
    for name in names:
        layers, Hq, Hkv, D = MODELS[name]
        for L in a.L:
            cap = L + R
            per_layer = 2 * a.B * Hkv * cap * D * 2
            n_rot = max(2, min(layers, (3 * 126 * 2**20) // per_layer + 1))  # rotate over > 3 x L2 of K/V
            g = torch.Generator(device=dev).manual_seed(1234)
            ks = [torch.randn(a.B, Hkv, cap, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(n_rot)]
            vs = [torch.randn(a.B, Hkv, cap, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(n_rot)]
            q = torch.randn(a.B, Hq, R, D, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
            out = torch.empty_like(q)
            lens = torch.full((a.B,), L, dtype=torch.int32, device=dev)
            ws = torch.empty(int(cuda_ops._capi.lib().sd_verify_attention_workspace_bytes(a.B, Hq, R, D, 0)),
                             dtype=torch.uint8, device=dev)
            bytes_alg = per_layer + 2 * a.B * Hq * R * D * 2
            for cfg in a.configs:
                cuda_ops.set_attention_tuning(cfg)
                for ns in a.splits:
                    try:
                        fns = [lambda i=i: cuda_ops.verify_attention_cuda(q, ks[i], vs[i], lens, max_seq_len=L, n_splits=ns,
                                                                          out=out, workspace=ws) for i in range(n_rot)]
                        us = time_replays(fns, a.iters)
                    except Exception as e:  # noqa: BLE001  (a split count the workspace cannot hold, ...)
                        print(json.dumps({"model": name, "L": L, "config": cfg, "splits": ns, "error": str(e)[:120]}), flush=True)
                        continue
                    gbs = bytes_alg / us / 1e3
                    print(json.dumps({"model": name, "D": D, "Hq": Hq, "Hkv": Hkv, "L": L, "config": cfg, "splits": ns,
                                      "us": round(us, 2), "GBps": round(gbs), "frac_of_measured_peak": round(gbs / peak, 3),
                                      "rotating_layers": n_rot}), flush=True)
            cuda_ops.set_attention_tuning(0)
            del ks, vs
            torch.cuda.empty_cache()
    return 0


if __name__ == "__main__":
    sys.exit(main())
