#!/usr/bin/env python3
"""Sweep launch configurations of the fused verify kernel (and kv_append variants) on one GPU.

    python tools/sweep_verify.py [--B 64] [--K 4] [--V 128256] [--iters 200]

Prints one line per configuration: microseconds per launch (CUDA events around graph replays over
rotating input buffers, so reads come from HBM, not the 126 MB L2) and achieved GB/s."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import torch  # noqa: E402

from kernels import cuda_ops  # noqa: E402
from specdec.cache.kv_arena import KVArena  # noqa: E402
from specdec.verify_step import VerifyStep  # noqa: E402


GROUP = True  # capture all rotating launches into ONE graph (PDL edges between them) vs one graph each


def time_graphs(fns, iters, warm=20):
    side = torch.cuda.Stream()
    graphs = []
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        if GROUP:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for fn in fns:
                    fn()
            graphs.append(g)
        else:
            for fn in fns:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    fn()
                graphs.append(g)
    torch.cuda.synchronize()
    per = len(fns) if GROUP else 1
    reps = max(1, iters // per)
    for i in range(max(2, warm // per)):
        graphs[i % len(graphs)].replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        graphs[i % len(graphs)].replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * per) * 1e3  # us per launch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--V", type=int, default=128256)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--nbuf", type=int, default=4)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--skip-kv", action="store_true")
    ap.add_argument("--skip-verify", action="store_true")
    ap.add_argument("--skip-attn", action="store_true")
    ap.add_argument("--L", type=int, default=1024)
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--single-graphs", action="store_true", help="one graph per launch (no PDL edges)")
    ap.add_argument("--quick", action="store_true", help="few configurations")
    ap.add_argument("--extras", action="store_true", help="sampling-mode verify and bonus sampler timings")
    ap.add_argument("--tma", action="store_true", help="only the TMA-ring verify variant sweep")
    ap.add_argument("--stats", action="store_true")
    args = ap.parse_args()
    global GROUP
    GROUP = not args.single_graphs
    cuda_ops.set_pdl(bool(args.pdl))
    print(json.dumps({"pdl": args.pdl, "group_graph": GROUP, "B": args.B, "K": args.K, "V": args.V}), flush=True)
    dev = torch.device("cuda")
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    B, K, V = args.B, args.K, args.V
    steps = []
    for i in range(args.nbuf):
        st = VerifyStep(B, K, V, arena=None, device=dev, logits_dtype=dt, mode="typical" if args.stats else "greedy",
                        threshold=0.3)
        st.logits.copy_(torch.randn(B, K + 1, V, device=dev).to(dt))
        st.draft_ids.copy_(torch.randint(0, V, (B, K), device=dev))
        steps.append(st)
    nbytes = B * (K + 1) * V * steps[0].logits.element_size()
    results = []
    if not args.skip_verify:
        # (variant, ctas_per_sm, threads, vpt): 1 = persistent equal-range grid, 0 = CTA per (row, chunk)
        configs = [(3, 0, 0, 0), (1, 0, 0, 0), (0, 0, 0, 0)]
        # variant 2 (TMA ring): vpt = vectors per thread per ring slot (2 -> 8 KB, 4 -> 16 KB), cps = CTAs per SM
        configs += [(2, c, 256, v) for v in (2, 4) for c in (1, 2, 3)]
        configs += [(3, 0, 256, v) for v in (4, 6, 7, 8, 9, 11, 12)]
        if args.tma:
            configs = [(0, 0, 0, 0)] + [(2, c, 256, v) for v in (2, 4) for c in (1, 2, 3, 4)] + [(3, 0, 256, v) for v in (4, 6, 7, 8, 9, 11, 12)]
        if not args.quick:
            configs += [(1, c, t, v) for t in (256, 512) for v in (2, 4) for c in (0, 4) if not (t == 512 and c == 4)]
            configs += [(0, 0, t, v) for t in (128, 256, 512) for v in (4, 5, 6, 7, 8, 16)]
        for (var, cps, t, v) in configs:
            try:
                cuda_ops.set_verify_variant(var, cps)
                cuda_ops.set_verify_tuning(t, v)
                us = time_graphs([s.run for s in steps], args.iters)
                results.append({"kernel": "verify", "variant": var, "ctas_per_sm": cps, "threads": t, "vpt": v, "us": us,
                                "GBps": nbytes / us / 1e3})
                print(json.dumps(results[-1]), flush=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"kernel": "verify", "variant": var, "threads": t, "vpt": v, "error": str(e)[:200]}), flush=True)
        cuda_ops.set_verify_variant()
    # torch.argmax on the same buffers: the ATen incumbent (reference.py:37 on CUDA tensors)
    us = time_graphs([lambda s=s: torch.argmax(s.logits, dim=-1) for s in steps], args.iters)
    print(json.dumps({"kernel": "torch.argmax", "us": us, "GBps": nbytes / us / 1e3}), flush=True)
    # copy peak on this box, same method as MEASURED_PEAKS (read+write bytes)
    src = [s.logits for s in steps]
    dst = [torch.empty_like(s.logits) for s in steps[:2]]
    us = time_graphs([lambda i=i: dst[i % 2].copy_(src[i]) for i in range(args.nbuf)], args.iters)
    print(json.dumps({"kernel": "torch.copy_", "us": us, "GBps": 2 * nbytes / us / 1e3}), flush=True)

    if not args.skip_kv:
        layers, H, D, R, L = 28, 8, 128, K + 1, 1024
        arena = KVArena(layers, B, H, D, L + R * (args.iters + 64) * 2, R, torch.bfloat16, dev)
        stages = [(torch.randn(arena.new_k.shape, device=dev).to(torch.bfloat16), torch.randn(arena.new_v.shape, device=dev).to(torch.bfloat16))
                  for _ in range(args.nbuf)]
        tables = [arena._build_table(k, v) for k, v in stages]
        accept = torch.randint(0, K + 1, (B,), device=dev, dtype=torch.int32)
        mean_keep = float((accept + 1).float().mean())
        for variant in ((1, 3) if args.quick else (1, 2, 3, 3 + 16 * 2, 3 + 16 * 3)):
            for zero in (True, False):
                cuda_ops.set_kv_variant(variant)
                arena.set_lengths([L] * B)

                def mk(i):
                    def f():
                        arena.layer_table = tables[i]
                        arena.append(accept, keep_bias=1, zero_rejected=zero, check_capacity=False)
                    return f

                us = time_graphs([mk(i) for i in range(args.nbuf)], args.iters)
                nb = arena.step_bytes(mean_keep, R, zero)
                print(json.dumps({"kernel": "kv_append", "variant": variant, "zero_rejected": zero, "us": us,
                                  "GBps": nb / us / 1e3, "bytes": nb, "mean_keep": mean_keep}), flush=True)
        cuda_ops.set_kv_variant(0)
        del arena, stages, tables

    if args.extras:
        # sampling-mode verify (softmax statistics in the same pass) and the bonus sampler
        for mode in ("typical", "confidence", "topk"):
            sts = []
            for i in range(args.nbuf):
                st = VerifyStep(B, K, V, arena=None, device=dev, logits_dtype=dt, mode=mode, threshold=5.0 if mode == "topk" else 0.3)
                st.logits.copy_(steps[i].logits)
                st.draft_ids.copy_(steps[i].draft_ids)
                sts.append(st)
            us = time_graphs([s.run for s in sts], args.iters)
            print(json.dumps({"kernel": "verify_step", "mode": mode, "us": us, "GBps": nbytes / us / 1e3}), flush=True)
            del sts
        acc = torch.randint(0, K + 1, (B,), device=dev, dtype=torch.int32)
        u = torch.rand(B, device=dev)
        out = torch.empty(B, dtype=torch.int32, device=dev)
        for name, kw in (("greedy", dict(uniform=None)), ("T0.7", dict(uniform=u, temperature=0.7)),
                         ("T0.7_k50_p0.9 (reference default)", dict(uniform=u, temperature=0.7, top_k=50, top_p=0.9)),
                         ("T0.7_p0.9", dict(uniform=u, temperature=0.7, top_p=0.9))):
            us = time_graphs([lambda s=s: cuda_ops.sample_bonus_cuda(s.logits, row_index=acc, out=out, **kw) for s in steps], args.iters)
            print(json.dumps({"kernel": "sample_bonus", "config": name, "us": us, "row_MB": B * V * steps[0].logits.element_size() / 1e6}), flush=True)
        # the reference's per-sequence sampler on the same GPU rows (torch ops), for scale
        def ref_sampler(s):
            for b in range(B):
                row = s.logits[b, int(0)] / 0.7
                tk, ti = torch.topk(row, 50)
                f = torch.full_like(row, float("-inf")).scatter_(-1, ti, tk)
                sl, si = torch.sort(f, descending=True)
                pr = torch.softmax(sl.float(), -1)
                rm = torch.cumsum(pr, -1) > 0.9
                rm[0] = False
                f = f.masked_fill(rm.scatter(-1, si, rm), float("-inf"))
                torch.multinomial(torch.softmax(f.float(), -1), 1)
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ref_sampler(steps[0])
        torch.cuda.synchronize()
        a.record()
        ref_sampler(steps[1])
        b2.record()
        torch.cuda.synchronize()
        print(json.dumps({"kernel": "torch per-sequence sampler (topk+sort+cumsum+multinomial x B)", "us": a.elapsed_time(b2) * 1e3}), flush=True)

    if not args.skip_attn:
        # one layer of verify attention per launch, rotating over `nl` layers of arena so KV comes from HBM
        for (name, Hq, Hkv, D) in (("llama3.2-3b", 24, 8, 128), ("llama3.2-1b", 32, 8, 64)):
            R, L, nl = K + 1, args.L, 6
            cap = L + R
            kk = torch.randn(nl, B, Hkv, cap, D, device=dev).to(torch.bfloat16)
            vv = torch.randn(nl, B, Hkv, cap, D, device=dev).to(torch.bfloat16)
            q = torch.randn(B, Hq, R, D, device=dev).to(torch.bfloat16)
            out = torch.empty_like(q)
            sl = torch.full((B,), L, dtype=torch.int32, device=dev)
            nb = 2 * B * Hkv * (L + R) * D * 2 + 2 * B * Hq * R * D * 2
            flops = 4 * B * Hq * R * (L + R) * D
            ws = torch.empty(int(cuda_ops._capi.lib().sd_verify_attention_workspace_bytes(B, Hq, R, D, 8)), dtype=torch.uint8, device=dev)
            for config in ((1,) if args.quick else (1, 2, 3)):
                for ns in ((0,) if args.quick else (0, 1, 2)):
                    cuda_ops.set_attention_tuning(config)
                    fns = [lambda l=l: cuda_ops.verify_attention_cuda(q, kk[l], vv[l], sl, max_seq_len=L, n_splits=ns, out=out, workspace=ws)
                           for l in range(nl)]
                    try:
                        us = time_graphs(fns, args.iters)
                        print(json.dumps({"kernel": "verify_attention", "model": name, "config": config, "n_splits": ns, "us": us,
                                          "GBps": nb / us / 1e3, "TFLOPs": flops / us / 1e6, "bytes": nb}), flush=True)
                    except Exception as e:  # noqa: BLE001
                        print(json.dumps({"kernel": "verify_attention", "model": name, "config": config, "n_splits": ns, "error": str(e)[:200]}), flush=True)
            cuda_ops.set_attention_tuning(0)
            # incumbent: torch SDPA with an explicit mask on the same data (one call for all K+1 queries)
            mask = torch.arange(cap, device=dev)[None, :] <= (L + torch.arange(R, device=dev))[:, None]
            fns = [lambda l=l: torch.nn.functional.scaled_dot_product_attention(q, kk[l], vv[l], attn_mask=mask, enable_gqa=True) for l in range(nl)]
            try:
                us = time_graphs(fns, args.iters)
                print(json.dumps({"kernel": "torch.sdpa(gqa,mask)", "model": name, "us": us, "GBps": nb / us / 1e3}), flush=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"kernel": "torch.sdpa", "model": name, "error": str(e)[:200]}), flush=True)
            del kk, vv


if __name__ == "__main__":
    main()
