#!/usr/bin/env python3
"""Single-GPU cost breakdown of the peer exchange (world = 1: the 'peer' is the GPU itself)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import torch  # noqa: E402

from scheduler.batch_shard import PeerExchange  # noqa: E402
from specdec.cache.kv_arena import KVArena  # noqa: E402
from specdec.verify_step import VerifyStep  # noqa: E402

dev = torch.device("cuda")
B, K, V, NB = 64, 4, 128256, 4
x = PeerExchange.emulate(1, B, K, dev)[0]
steps = []
for i in range(NB):
    st = VerifyStep(B, K, V, device=dev, exchange=x)
    st.exchange_wait = False
    st.logits.copy_(torch.randn(B, K + 1, V, device=dev).to(torch.bfloat16))
    st.draft_ids.copy_(torch.randint(0, V, (B, K), device=dev))
    steps.append(st)
arena = KVArena(28, B, 8, 128, 1024 + 5 * 64, K + 1, torch.bfloat16, dev)
lens0 = torch.full((B,), 1024, dtype=torch.int32, device=dev)
side = torch.cuda.Stream()


def timed(fn, n=400):
    for i in range(NB):
        fn(i)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(NB):
                fn(i)
    torch.cuda.synchronize()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n // NB):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def set_x(on):
    for st in steps:
        st.exchange = x if on else None


def verify(i):
    steps[i].run()


def verify_wait(i):
    steps[i].run()
    x.wait()


def step_plain(i):
    steps[i].run()
    arena.seq_lens.copy_(lens0)
    arena.append(steps[i].accept_len, keep_bias=1, check_capacity=False)


def step_fused(i):
    steps[i].run()
    arena.seq_lens.copy_(lens0)
    arena.append(steps[i].accept_len, keep_bias=1, check_capacity=False, exchange=x)


out = {}
set_x(False)
out["verify"] = timed(verify)
out["step_plain"] = timed(step_plain)
set_x(True)
out["verify_push_only"] = timed(verify)  # records pushed and counted, never consumed (timing only)
torch.cuda.synchronize()
# re-synchronise the counters with the epoch: a fresh exchange
x.close()
x = PeerExchange.emulate(1, B, K, dev)[0]
set_x(True)
out["verify_push_wait_kernel"] = timed(verify_wait)
out["step_fused_append"] = timed(step_fused)
print(json.dumps(out))
