#!/usr/bin/env python3
"""GPU check of the EXPERIMENTAL cluster sampler against the product sampler (needs one B200).

    python experiments/sample_cluster/run_gpu.py            # parity (bit-identical tokens) + timing
    python experiments/sample_cluster/run_gpu.py --no-time

Parity criterion: for every case, tokens and kept-set sizes equal those of sd_sample_bonus (which is itself
checked against the oracle in tests/test_sample_gpu.py).  Exit code 0 only if every case matches.
"""
import argparse
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import torch  # noqa: E402

from kernels import cuda_ops  # noqa: E402

LIB = HERE / "libsd_sample_cluster.so"
DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def build() -> None:
    src = HERE / "sd_sample_cluster.cu"
    if LIB.exists() and LIB.stat().st_mtime >= src.stat().st_mtime:
        return
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "-shared", "-I", str(ROOT / "include"), "-o", str(LIB), str(src)]
    subprocess.run(cmd, check=True)


def load():
    lib = C.CDLL(str(LIB))
    f = lib.sd_sample_bonus_cluster
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int,
                  C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return f


def cluster_sample(f, logits, row_index, uniform, temperature, top_k, top_p, cluster, out, kept):
    B, rows, V = logits.shape
    rc = f(logits.data_ptr(), DT[logits.dtype], logits.stride(0), logits.stride(1),
           row_index.data_ptr() if row_index is not None else None, 0, 0,
           uniform.data_ptr() if uniform is not None else None, float(temperature), int(top_k or 0),
           float(top_p if top_p is not None else 1.0), out.data_ptr(), kept.data_ptr(), B, rows, V, cluster,
           torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"sd_sample_bonus_cluster: {rc}")


def fits(V, cluster, dtype) -> bool:
    per = (V + cluster - 1) // cluster
    slice_ = (per + 7) & ~7
    return 32 * 256 * 12 + 1024 * 8 + slice_ * (4 if dtype == torch.float32 else 2) + 10240 <= 227 * 1024


def make_logits(B, rows, V, dtype, kind, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, rows, V, generator=g, device="cuda", dtype=torch.float32) * 3.0
    if kind == "ties":  # few distinct values: many ties at the top-k threshold, ranked by index
        x = torch.round(x * 2.0) / 2.0
    elif kind == "flat":  # every element equal: everything is a tie
        x = torch.zeros_like(x)
    elif kind == "peaked":
        x[:, :, 7] += 40.0
    return x.to(dtype)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-time", action="store_true")
    ap.add_argument("--quick", action="store_true", help="a handful of bf16 cases only")
    ap.add_argument("--time-only", action="store_true")
    a = ap.parse_args()
    build()
    f = load()
    cases = []
    for dtype in (torch.bfloat16, torch.float16, torch.float32):
        for (B, V) in ((64, 128256), (7, 50257), (5, 1000), (3, 37)):
            for kind in ("randn", "ties", "flat", "peaked"):
                for (T, k, p_) in ((0.7, 50, 0.9), (1.0, 1, 1.0), (1.3, 1024, 0.5), (0.7, 17, 1.0), (None, None, None),
                                   (1.0, None, 0.9), (0.8, 5000, 0.95), (1.0, None, None), (0.5, 2000, None)):  # full-row path
                    cases.append((dtype, B, V, kind, T, k, p_))
    if a.quick:
        cases = [(torch.bfloat16, B, V, kind, T, k, p_) for (B, V) in ((64, 128256), (5, 1000)) for kind in ("randn", "ties")
                 for (T, k, p_) in ((0.7, 50, 0.9), (None, None, None), (1.0, None, 0.9))]
    bad = 0
    ran = 0
    if a.time_only:
        cases = []
    for ci, (dtype, B, V, kind, T, k, p_) in enumerate(cases):
        if k is not None and k >= V:
            continue
        rows = 5
        logits = make_logits(B, rows, V, dtype, kind, 1000 + ci)
        row_index = torch.randint(0, rows, (B,), device="cuda", dtype=torch.int32)
        uniform = torch.rand(B, device="cuda") if T is not None else None
        want = torch.empty(B, dtype=torch.int32, device="cuda")
        want_kept = torch.zeros(B, dtype=torch.int32, device="cuda")
        cuda_ops.sample_bonus_cuda(logits, row_index=row_index, uniform=uniform, temperature=T or 1.0, top_k=k, top_p=p_,
                                   out=want, kept_count=want_kept)
        for cluster in ((2, 8) if a.quick else (1, 2, 4, 8)):
            if not fits(V, cluster, dtype):
                continue  # the slice does not fit in shared memory next to the scratch
            got = torch.full((B,), -5, dtype=torch.int32, device="cuda")
            got_kept = torch.full((B,), -5, dtype=torch.int32, device="cuda")
            cluster_sample(f, logits, row_index, uniform, T or 1.0, k, p_, cluster, got, got_kept)
            torch.cuda.synchronize()
            ran += 1
            if not (torch.equal(got, want) and torch.equal(got_kept, want_kept)):
                bad += 1
                nd = int((got != want).sum())
                print(json.dumps({"mismatch": [str(dtype), B, V, kind, T, k, p_], "cluster": cluster, "tokens_differ": nd,
                                  "kept_differ": int((got_kept != want_kept).sum())}), flush=True)
    print(json.dumps({"cases_run": ran, "mismatching": bad}), flush=True)
    if bad or a.no_time:
        return 1 if bad else 0

    # ---- timing: reference defaults (T=0.7, k=50, p=0.9), B=64, V=128256 bf16, rows selected by accept_len ----
    B, rows, V = 64, 5, 128256
    sets = [make_logits(B, rows, V, torch.bfloat16, "randn", 77 + i) for i in range(4)]  # 4 x 82 MB > L2
    row_index = torch.randint(0, rows, (B,), device="cuda", dtype=torch.int32)
    uniform = torch.rand(B, device="cuda")
    out = torch.empty(B, dtype=torch.int32, device="cuda")
    kept = torch.zeros(B, dtype=torch.int32, device="cuda")

    def timed(fn, n=200):
        for i in range(8):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    res = {"product_us": timed(lambda i: cuda_ops.sample_bonus_cuda(sets[i % 4], row_index=row_index, uniform=uniform,
                                                                    temperature=0.7, top_k=50, top_p=0.9, out=out))}
    for cluster in (2, 4, 8):
        if not fits(V, cluster, torch.bfloat16):
            continue
        res[f"cluster{cluster}_us"] = timed(lambda i: cluster_sample(f, sets[i % 4], row_index, uniform, 0.7, 50, 0.9, cluster, out, kept))
    res["row_bytes_total"] = B * V * 2
    print(json.dumps(res), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
