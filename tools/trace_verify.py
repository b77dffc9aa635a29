#!/usr/bin/env python3
"""Per-CTA timeline of the fused verify kernel (variant 0) from %globaltimer stamps.

    python tools/trace_verify.py [--B 64] [--K 4] [--V 128256] [--tuning 256,8]

Prints when CTAs start / finish scanning / exit relative to the first CTA's entry, per wave, and what
is left after the last byte has been scanned (the finalisation tail)."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kernels import _capi, cuda_ops  # noqa: E402
from specdec.verify_step import VerifyStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--V", type=int, default=128256)
    ap.add_argument("--tuning", default="0,0")
    ap.add_argument("--nbuf", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda")
    t, v = (int(x) for x in a.tuning.split(","))
    cuda_ops.set_verify_variant(0, 0)
    cuda_ops.set_verify_tuning(t, v)
    steps = []
    for i in range(a.nbuf):
        st = VerifyStep(a.B, a.K, a.V, device=dev)
        st.logits.copy_(torch.randn(a.B, a.K + 1, a.V, device=dev).to(torch.bfloat16))
        st.draft_ids.copy_(torch.randint(0, a.V, (a.B, a.K), device=dev))
        steps.append(st)
    cap = 1 << 16
    buf = torch.zeros(cap * 4, dtype=torch.int64, device=dev)
    for s in steps:  # warm, untraced
        s.run()
    torch.cuda.synchronize()
    res = []
    for rep in range(6):
        buf.zero_()
        _capi.lib().sd_verify_set_trace(buf.data_ptr(), cap)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps[rep % a.nbuf].run()
        e1.record()
        torch.cuda.synchronize()
        _capi.lib().sd_verify_set_trace(None, 0)
        tr = buf.cpu().numpy().reshape(cap, 4)
        tr = tr[tr[:, 0] != 0]
        t0 = tr[:, 0].min()
        start, scan, end, sm = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3, tr[:, 3]
        order = np.argsort(start)
        n = len(tr)
        out = {
            "rep": rep, "ctas": int(n), "sms": int(len(np.unique(sm))), "event_us": e0.elapsed_time(e1) * 1e3,
            "first_to_last_exit_us": float(end.max()),
            "start_us_pct": [float(np.percentile(start, q)) for q in (0, 25, 50, 75, 90, 99, 100)],
            "scan_done_us_pct": [float(np.percentile(scan, q)) for q in (0, 25, 50, 75, 90, 99, 100)],
            "cta_scan_duration_us_pct": [float(np.percentile(scan - start, q)) for q in (0, 50, 90, 100)],
            "cta_finish_duration_us_pct": [float(np.percentile(end - scan, q)) for q in (0, 50, 90, 99, 100)],
            "tail_after_last_scan_us": float(end.max() - scan.max()),
            "ctas_started_by_us": {str(x): int((start <= x).sum()) for x in (0.5, 1, 2, 4, 6, 8, 10, 12, 14)},
        }
        res.append(out)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
