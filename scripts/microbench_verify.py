#!/usr/bin/env python3
"""Microbenchmark of the registry ops against the ON-BOX INCUMBENTS -- the counterpart of the reference's
scripts/microbench_verify.py (same planted-match protocol, :57-64, its fp32 shapes :35-49) extended to the
shapes of BASELINE configs[1]: B = 1..256, K = 1..8, V = 50257, bf16, and to kv_append and the verify attention.

    python scripts/microbench_verify.py [--quick] [--sections verify,kv,attn] [--markdown profiles/r2_microbench.md]

verify   kernels.verify_prefix (this repo)  vs  (a) the reference's Triton kernel (src/kernels/triton/verify.py:101-169,
         loaded unmodified from baseline/_ref when that copy and Triton are present: the kernel the reference registers
         for CUDA)  and (b) torch.argmax + the host prefix loop (== verify_prefix_ref, reference.py:37-54, on CUDA tensors).
kv       kernels.kv_append in a `cache = kv_append(cache, chunk)` loop (this repo: O(Kn) bytes into an arena) and
         KVArena.append (all layers, one launch)  vs  torch.cat (kv_append_ref, reference.py:59-93: O(L) bytes per call),
         three model geometries, B <= 256, Kn <= 8.
attn     cuda_ops.verify_attention_cuda (K+1 queries, one pass)  vs  torch SDPA (GQA via enable_gqa, explicit mask) on the
         same tensors.
Times are wall-clock per call with a synchronise on both sides, as in the reference script (so they include launch
latency and any host loop -- that is the point of a drop-in comparison); identical results are asserted.
Nothing is imported from oracle/."""
import argparse
import importlib.util
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kernels import cuda_ops, get_kernel_info, kv_append, verify_prefix  # noqa: E402
from specdec.cache.kv_arena import KVArena  # noqa: E402

REF_TRITON = ROOT / "baseline" / "_ref" / "src" / "kernels" / "triton" / "verify.py"


def load_triton_incumbent():
    """The reference's verify_prefix_triton, unmodified, or (None, why)."""
    if not REF_TRITON.exists():
        return None, "baseline/_ref missing"
    try:
        import triton  # noqa: F401

        spec = importlib.util.spec_from_file_location("_ref_triton_verify", REF_TRITON)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.verify_prefix_triton, None
    except Exception as e:  # noqa: BLE001
        return None, f"{type(e).__name__}: {e}"[:160]


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


def emit(rows, rec):
    rows.append(rec)
    print(json.dumps(rec), flush=True)


def section_verify(a, rows):
    triton_fn, why = load_triton_incumbent()
    print(json.dumps({"section": "verify", "triton_incumbent": "loaded from baseline/_ref" if triton_fn else f"unavailable ({why})"}))
    rng = np.random.default_rng(1234)
    configs = [(1, 1, 4096, "fp32"), (1, 4, 4096, "fp32"), (8, 4, 4096, "fp32"), (1, 4, 32768, "fp32")]  # reference shapes
    bs = (1, 8, 64, 256) if a.quick else (1, 2, 4, 8, 16, 32, 64, 128, 256)
    ks = (1, 4, 8) if a.quick else (1, 2, 4, 8)
    configs += [(B, K, 50257, "bf16") for B in bs for K in ks]
    configs += [(64, 4, 128256, "bf16"), (1, 4, 128256, "bf16")]
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
        rec = {"section": "verify", "B": B, "K": K, "V": V, "dtype": dt,
               "kernel_us": round(timeit(lambda: verify_prefix(logits, ids), 50), 1),
               "torch_incumbent_us": round(timeit(lambda: incumbent(logits, ids), 5 if B * K > 256 else 20), 1)}
        if triton_fn is not None:
            try:
                for _ in range(3):
                    t_out = triton_fn(logits, ids)
                rec["triton_incumbent_us"] = round(timeit(lambda: triton_fn(logits, ids), 5 if B * K * V > 5e7 else 20), 1)
                rec["triton_identical"] = bool(torch.equal(t_out[0], got[0]))
                # the reference calls its kernel once PER SEQUENCE (policies.py:128-134 from pipeline.py:3018): B launches per step
                if B > 1 and B <= 64:
                    rec["triton_per_sequence_calls_us"] = round(timeit(lambda: [triton_fn(logits[b:b + 1], ids[b:b + 1]) for b in range(B)], 3), 1)
                    rec["kernel_per_sequence_calls_us"] = round(timeit(lambda: [verify_prefix(logits[b:b + 1], ids[b:b + 1]) for b in range(B)], 5), 1)
            except Exception as e:  # noqa: BLE001
                rec["triton_incumbent_us"] = None
                rec["triton_error"] = f"{type(e).__name__}: {e}"[:120]
        rec.update({"speedup_vs_torch": round(rec["torch_incumbent_us"] / rec["kernel_us"], 1), "identical": ok,
                    "logits_MB": round(B * K * V * logits.element_size() / 1e6, 2)})
        if rec.get("triton_incumbent_us"):
            rec["speedup_vs_triton"] = round(rec["triton_incumbent_us"] / rec["kernel_us"], 1)
        emit(rows, rec)


GEOMS = {"llama-3.2-3b": (28, 8, 128), "llama-3.2-1b": (16, 8, 64), "gpt2": (12, 12, 64)}  # layers, KV heads, head dim


def section_kv(a, rows):
    print(json.dumps({"section": "kv"}))
    L = 1024
    for name, (layers, H, D) in GEOMS.items():
        for B in ((8, 64) if a.quick else (1, 8, 64, 256)):
            for Kn in ((4,) if a.quick else (1, 4, 8)):
                if B * layers * H * (L + 64) * D * 2 * 4 > 60e9:
                    continue
                g = torch.Generator(device="cuda").manual_seed(1)
                base = [(torch.randn(B, H, L, D, device="cuda", generator=g).to(torch.bfloat16),
                         torch.randn(B, H, L, D, device="cuda", generator=g).to(torch.bfloat16)) for _ in range(layers)]
                new = [(torch.randn(B, H, Kn, D, device="cuda", generator=g).to(torch.bfloat16),
                        torch.randn(B, H, Kn, D, device="cuda", generator=g).to(torch.bfloat16)) for _ in range(layers)]

                def cat_all():  # kv_append_ref per layer (hf_wrappers.py:985-1030 calls it once per layer)
                    return [(torch.cat([bk, nk], 2), torch.cat([bv, nv], 2)) for (bk, bv), (nk, nv) in zip(base, new)]

                t_cat = timeit(cat_all, 5)
                # drop-in: cache = kv_append(cache, chunk); the first call adopts the base (one copy), then O(Kn) per call
                cache = [kv_append(bk, bv, nk, nv) for (bk, bv), (nk, nv) in zip(base, new)]
                want = cat_all()
                same = all(torch.equal(c[0], w[0]) and torch.equal(c[1], w[1]) for c, w in zip(cache, want))
                state = {"c": cache}

                def dropin():
                    state["c"] = [kv_append(ck, cv, nk, nv) for (ck, cv), (nk, nv) in zip(state["c"], new)]

                t_drop = timeit(dropin, 10)
                del cache, state, want
                arena = KVArena(layers, B, H, D, L + 64 * Kn + 16, Kn, torch.bfloat16, "cuda")
                arena.set_lengths([L] * B)
                n_keep = torch.full((B,), Kn, dtype=torch.int32, device="cuda")
                t_arena = timeit(lambda: arena.append(n_keep, check_capacity=False), 20)
                del arena
                emit(rows, {"section": "kv", "model": name, "layers": layers, "B": B, "Kn": Kn, "L": L,
                            "torch_cat_us": round(t_cat, 1), "kv_append_dropin_us": round(t_drop, 1), "arena_append_us": round(t_arena, 1),
                            "speedup_dropin": round(t_cat / t_drop, 1), "speedup_arena": round(t_cat / t_arena, 1), "identical": same,
                            "cat_bytes_MB": round(layers * 2 * B * H * (2 * L + Kn) * D * 2 / 1e6, 1),
                            "inplace_bytes_MB": round(layers * 2 * 2 * B * H * Kn * D * 2 / 1e6, 2)})
                del base, new
                torch.cuda.empty_cache()


def section_attn(a, rows):
    print(json.dumps({"section": "attn"}))
    import torch.nn.functional as F

    R = 5
    for name, (Hq, Hkv, D) in {"llama-3.2-3b": (24, 8, 128), "llama-3.2-1b": (32, 8, 64), "gpt2": (12, 12, 64)}.items():
        for B, L in (((64, 1024),) if a.quick else ((8, 1024), (64, 1024), (64, 4096))):
            g = torch.Generator(device="cuda").manual_seed(2)
            q = torch.randn(B, Hq, R, D, device="cuda", generator=g).to(torch.bfloat16)
            k = torch.randn(B, Hkv, L + R, D, device="cuda", generator=g).to(torch.bfloat16)
            v = torch.randn(B, Hkv, L + R, D, device="cuda", generator=g).to(torch.bfloat16)
            lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
            mask = (torch.arange(L + R, device="cuda")[None, :] <= (L + torch.arange(R, device="cuda"))[:, None])

            def sdpa():
                return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, enable_gqa=Hq != Hkv)

            ours = cuda_ops.verify_attention_cuda(q, k, v, lens, max_seq_len=L)
            ref = sdpa()
            err = (ours.float() - ref.float()).abs().max().item()
            t_o = timeit(lambda: cuda_ops.verify_attention_cuda(q, k, v, lens, max_seq_len=L), 20)
            t_s = timeit(sdpa, 10)
            by = 2 * B * Hkv * (L + R) * D * 2 + 2 * B * Hq * R * D * 2
            emit(rows, {"section": "attn", "model": name, "B": B, "L": L, "Hq": Hq, "Hkv": Hkv, "D": D, "verify_attention_us": round(t_o, 1),
                        "torch_sdpa_us": round(t_s, 1), "speedup": round(t_s / t_o, 1), "max_abs_diff_vs_sdpa": round(err, 4),
                        "GBps": round(by / t_o / 1e3, 0)})


def write_markdown(path, rows, info):
    out = [f"# Registry ops vs the on-box incumbents ({info['device']})", "",
           "`python scripts/microbench_verify.py` -- wall clock per call with a synchronise on both sides (launch latency and host",
           "loops included, as in the reference's own microbenchmark); results asserted identical.", ""]
    for sec, title in (("verify", "verify_prefix"), ("kv", "kv_append (per step, all layers)"), ("attn", "verify attention (one layer)")):
        rs = [r for r in rows if r.get("section") == sec]
        if not rs:
            continue
        keys = [k for k in rs[0] if k != "section"]
        for r in rs:
            for k in r:
                if k != "section" and k not in keys:
                    keys.append(k)
        out += [f"## {title}", "", "| " + " | ".join(keys) + " |", "|" + "---|" * len(keys)]
        for r in rs:
            out.append("| " + " | ".join(" ".join(str(r.get(k, "")).split()).replace("|", "/") for k in keys) + " |")
        out.append("")
    Path(path).write_text("\n".join(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--sections", default="verify,kv,attn")
    ap.add_argument("--markdown", default="")
    a = ap.parse_args()
    info = get_kernel_info()
    info = {"verify_backend": info["verify_backend"], "kv_append_backend": info["kv_append_backend"], "device": torch.cuda.get_device_name(0)}
    print(json.dumps(info))
    rows = []
    secs = a.sections.split(",")
    if "verify" in secs:
        section_verify(a, rows)
    if "kv" in secs:
        section_kv(a, rows)
    if "attn" in secs:
        section_attn(a, rows)
    if a.markdown:
        write_markdown(a.markdown, rows, info)


if __name__ == "__main__":
    main()
