#!/usr/bin/env python3
"""Per-CTA timeline of the fused step kernel (verify_step_kernel, lagged append) from %globaltimer stamps.

    python tools/trace_step.py [--B 64] [--K 4] [--V 128256] [--lag 0] [--layers 28]
    torchrun --nproc-per-node N tools/trace_step.py --exchange [--xlag 1]   (the step with the peer-memory exchange;
                                                        --loopback on one GPU = the exchange against the GPU itself)

Prints, relative to the first CTA's entry: when the verify CTAs start / finish their scan / exit, when the tail
CTAs start / store / exit, and what is left after the last logits byte has been reduced (the tail of the launch).
Evidence tool: the traced launch is eager (no graph) and slightly slower than the benchmarked one."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kernels import _capi, cuda_ops  # noqa: E402
from specdec.cache.kv_arena import KVArena  # noqa: E402
from specdec.verify_step import VerifyStep  # noqa: E402


def pct(x, qs=(0, 25, 50, 75, 90, 99, 100)):
    return [round(float(np.percentile(x, q)), 2) for q in qs] if len(x) else []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--V", type=int, default=128256)
    ap.add_argument("--layers", type=int, default=28)
    ap.add_argument("--H", type=int, default=8)
    ap.add_argument("--D", type=int, default=128)
    ap.add_argument("--lag", type=int, default=0)
    ap.add_argument("--nbuf", type=int, default=4)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--exchange", action="store_true", help="under torchrun: PeerExchange between the ranks (lock-step at --xlag 0)")
    ap.add_argument("--loopback", action="store_true", help="one GPU: exchange records with the GPU itself")
    ap.add_argument("--xlag", type=int, default=0, help="0 = a step gathers its own records, 1 = the previous step's")
    a = ap.parse_args()
    import os
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    peer = None
    if a.exchange and world > 1:
        import torch.distributed as dist
        from scheduler.batch_shard import BatchShard, PeerExchange
        dist.init_process_group("nccl", device_id=dev)
        peer = PeerExchange(BatchShard(a.B * world, world, rank), a.K, dev, lag=a.xlag)
    elif a.loopback:
        from scheduler.batch_shard import PeerExchange
        peer = PeerExchange.emulate(1, a.B, a.K, dev, lag=a.xlag)[0]
    torch.manual_seed(1234 + rank)
    R = a.K + 1
    if a.lag:
        cuda_ops.set_fused_mode(0, a.lag)
    arena = KVArena(a.layers, a.B, a.H, a.D, 1024 + R * (a.reps + a.nbuf + 8), R, torch.bfloat16, dev)
    arena.set_lengths([1024] * a.B)
    steps, tables = [], []
    for i in range(a.nbuf):
        st = VerifyStep(a.B, a.K, a.V, device=dev)
        lg = torch.randn(a.B, R, a.V, device=dev).to(torch.bfloat16)
        ids = torch.randint(0, a.V, (a.B, a.K), device=dev)
        plant = torch.rand(a.B, a.K, device=dev) < 0.8
        bi, ki = torch.nonzero(plant, as_tuple=True)
        lg[bi, ki, ids[bi, ki]] = 10.0
        st.logits.copy_(lg)
        st.draft_ids.copy_(ids)
        steps.append(st)
        nk = torch.randn(arena.new_k.shape, device=dev).to(torch.bfloat16)
        nv = torch.randn(arena.new_v.shape, device=dev).to(torch.bfloat16)
        tables.append((arena._build_table(nk, nv), nk, nv))

    def run(i):
        st = steps[i]
        arena.use_staging(tables[i][1], tables[i][2], tables[i][0])
        cuda_ops.verify_append_launch(st.logits, st.draft_ids, st.accept_len, st.accepted_mask, st.target_ids, st.bonus,
                                      st.packed, st.workspace, st.seq_ready, arena, exchange=peer)

    for i in range(a.nbuf):
        run(i)
    torch.cuda.synchronize()
    cap = 1 << 16
    buf = torch.zeros(cap * 4, dtype=torch.int64, device=dev)
    for rep in range(a.reps):
        buf.zero_()
        _capi.lib().sd_verify_set_trace(buf.data_ptr(), cap)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if peer is not None and world > 1:
            # three untraced steps first: the lock-step exchange itself brings the ranks within a few us of each other
            _capi.lib().sd_verify_set_trace(None, 0)
            dist.barrier()
            torch.cuda.synchronize()
            for w in range(3):
                run((rep + 1 + w) % a.nbuf)
            _capi.lib().sd_verify_set_trace(buf.data_ptr(), cap)
        e0.record()
        run(rep % a.nbuf)
        e1.record()
        torch.cuda.synchronize()
        _capi.lib().sd_verify_set_trace(None, 0)
        tr = buf.cpu().numpy().reshape(cap, 4)
        n_all = int((tr[:, 0] != 0).sum())
        tr = tr[:n_all]
        t0 = tr[:, 0].min()
        start, mid, end = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3
        kind = (tr[:, 3] >> 16) & 3                   # 0 big verify (first `lag` sequences), 1 carrying verify, 2 tail
        n_verify = int((kind < 2).sum())
        n_big = int((kind == 0).sum())
        S = 0
        v = slice(0, n_verify)
        t = slice(n_verify, n_all)
        per_seq_end = np.array([end[v].min(), end[v].max()])
        out = {
            "rank": rank, "world": world, "exchange": None if peer is None else ("loopback" if world == 1 else f"peer lag {a.xlag}"),
            # with an exchange the first tail CTA carries no rows: it gathers the records and advances the exchange epoch
            "first_tail_cta_start_mid_exit_us": [round(float(x[n_verify]), 2) for x in (start, mid, end)] if n_all > n_verify else [],
            "rep": rep, "ctas": n_all, "verify_ctas": n_verify, "big_chunk_ctas": n_big, "tail_ctas": n_all - n_verify,
            "big_exit_pct": pct(end[:n_big]) if n_big else [],
            "event_us": round(e0.elapsed_time(e1) * 1e3, 2),
            "first_entry_to_last_exit_us": round(float(end.max()), 2),
            "verify_start_pct": pct(start[v]), "verify_scan_done_pct": pct(mid[v]), "verify_exit_pct": pct(end[v]),
            "verify_cta_life_pct": pct(end[v] - start[v], (0, 50, 90, 99, 100)),
            "verify_cta_after_scan_pct": pct(end[v] - mid[v], (0, 50, 90, 99, 100)),
            "tail_start_pct": pct(start[t]), "tail_stored_pct": pct(mid[t]), "tail_exit_pct": pct(end[t]),
            "tail_cta_life_pct": pct(end[t] - start[t], (0, 50, 90, 100)),
            "last_scan_done_us": round(float(mid[v].max()), 2),
            "after_last_scan_us": round(float(end.max() - mid[v].max()), 2),
            "seq_done_us_first_last": [round(float(per_seq_end[0]), 2), round(float(per_seq_end[-1]), 2)],
        }
        print(json.dumps(out), flush=True)
    assert cuda_ops.device_error() == 0


if __name__ == "__main__":
    main()
