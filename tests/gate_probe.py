#!/usr/bin/env python
"""Diagnostic for the start gate between emulated ranks (tests/test_scheduler_gpu.py:
test_start_gate_rendezvous_of_emulated_ranks): repeats the test's protocol on fresh exchange buffers and dumps the
exchange headers, the descriptors and sd_device_error when a rank passes its gate early.  Not collected by pytest.

    python tests/gate_probe.py [--trials 6]                  fresh process, with / without a GB-sized allocation first
    python tests/gate_probe.py --after-residency             after the B = 256 fused-step test (seconds of CPU-side oracle
                                                             work = seconds of GPU idle) -- the order that failed
    python tests/gate_probe.py --idle 3                      after 3 s of plain GPU idle (does idle alone reproduce it?)

What it found (profiles/r2j_gate_probe.md): the early pass was the bounded wait giving up -- sd_device_error = 16
(kSpinStartGate) with clean headers -- because two %globaltimer reads 40 us apart differed by more than the 2 s budget
on the first gate launch after the idle period.  csrc/sd_exchange.cu: GateSpinGuard.
"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "llm-inference-lab_b200"))
from kernels import cuda_ops  # noqa: E402
from scheduler.batch_shard import PeerExchange  # noqa: E402


def header(ptr):
    rt = C.CDLL("libcudart.so.12")
    buf = (C.c_uint64 * 32)()
    rc = rt.cudaMemcpy(buf, C.c_void_p(ptr), 256, 2)
    w = list(buf)
    return {"rc": rc, "epoch": w[0] & 0xFFFFFFFF, "gate": w[2] & 0xFFFFFFFF, "arrival": w[8:12]}


def raw_bytes(ptr, n):
    rt = C.CDLL("libcudart.so.12")
    buf = (C.c_uint8 * n)()
    rt.cudaMemcpy(buf, C.c_void_p(ptr), n, 2)
    return bytes(buf).hex()


def strict_trial(trial):
    """The test's own protocol, nothing added before the first gate launch (no header read, no synchronize); the
    completion of rank 0's event is polled with time stamps instead of one sleep."""
    xs = PeerExchange.emulate(2, 8, 4, "cuda")
    rec = {"strict_trial": trial, "ptrs": [hex(p) for p in xs[0].peers], "rounds": []}
    try:
        gates = [cuda_ops.StartGate("cuda:0", x) for x in xs]
        streams = [torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()]
        flags = [torch.zeros(1, device="cuda") for _ in range(2)]
        peek = torch.zeros(1).pin_memory()
        for rnd in range(3):
            r = {"rnd": rnd}
            evs = []
            with torch.cuda.stream(streams[0]):
                gates[0].arm(hold=True)
                flags[0].add_(1.0)
                evs.append(torch.cuda.Event())
                evs[0].record()
            t_rel = time.perf_counter()
            gates[0].release()
            early = None
            while time.perf_counter() - t_rel < 0.02:
                if evs[0].query():
                    early = round((time.perf_counter() - t_rel) * 1e6, 1)
                    break
            r["r0_passed_after_us"] = early
            if early is not None:
                with torch.cuda.stream(streams[2]):
                    peek.copy_(flags[0], non_blocking=True)
                streams[2].synchronize()
                r["flag0_seen_by_side_stream"] = float(peek[0])
                r["hdr0"], r["hdr1"] = header(xs[0].peers[0]), header(xs[0].peers[1])
                r["desc0"], r["desc1"] = raw_bytes(xs[0].desc_ptr, 96), raw_bytes(xs[1].desc_ptr, 96)
                r["host_flags"] = [int(g.flag[0]) for g in gates]
            with torch.cuda.stream(streams[1]):
                gates[1].arm(hold=True)
                flags[1].add_(1.0)
                evs.append(torch.cuda.Event())
                evs[1].record()
            time.sleep(0.01)
            r["done_before_r1_release"] = [evs[0].query(), evs[1].query()]
            gates[1].release()
            t0 = time.perf_counter()
            while not (evs[0].query() and evs[1].query()):
                if time.perf_counter() - t0 > 3.0:
                    r["never_opened"] = True
                    break
            r["open_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
            rec["rounds"].append(r)
        torch.cuda.synchronize()
        rec["device_error"] = cuda_ops.device_error()
    finally:
        for x in xs:
            x.close()
    return rec


def one_trial(trial, dirty):
    if dirty:  # what the B = 256 fused-step test leaves behind: GB-sized blocks in torch's allocator, a busy null stream
        g = torch.Generator(device="cuda").manual_seed(trial)
        big = torch.randn(256, 5, 128256, generator=g, device="cuda").to(torch.bfloat16)
        big.add_(1)
        del big
    xs = PeerExchange.emulate(2, 8, 4, "cuda")
    rec = {"trial": trial, "dirty": dirty, "ptrs": [hex(p) for p in xs[0].peers], "hdr0_at_start": header(xs[0].peers[0]),
           "hdr1_at_start": header(xs[0].peers[1]), "rounds": []}
    try:
        gates = [cuda_ops.StartGate("cuda:0", x) for x in xs]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        for rnd in range(3):
            r = {"rnd": rnd}
            evs = [torch.cuda.Event(), torch.cuda.Event()]
            with torch.cuda.stream(streams[0]):
                gates[0].arm(hold=True)
                evs[0].record()
            time.sleep(0.005)
            r["r0_done_before_release"] = evs[0].query()
            gates[0].release()
            time.sleep(0.02)
            r["r0_done_before_r1_armed"] = evs[0].query()
            if r["r0_done_before_r1_armed"]:
                r["hdr0"], r["hdr1"] = header(xs[0].peers[0]), header(xs[0].peers[1])
            with torch.cuda.stream(streams[1]):
                gates[1].arm(hold=True)
                evs[1].record()
            time.sleep(0.02)
            r["done_before_r1_release"] = [evs[0].query(), evs[1].query()]
            gates[1].release()
            t0 = time.perf_counter()
            while not (evs[0].query() and evs[1].query()):
                if time.perf_counter() - t0 > 3.0:
                    r["never_opened"] = True
                    break
            r["open_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
            rec["rounds"].append(r)
        torch.cuda.synchronize()
        rec["device_error"] = cuda_ops.device_error()
    finally:
        for x in xs:
            x.close()
    rec["bad"] = any(r["r0_done_before_release"] or r["r0_done_before_r1_armed"] or any(r["done_before_r1_release"]) for r in rec["rounds"])
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=6)
    ap.add_argument("--after-residency", action="store_true",
                    help="first run tests/test_verify_step_gpu.py::test_fused_step_with_a_grid_larger_than_residency[0] in "
                         "this process (the order in which the rendezvous test failed), then the strict protocol")
    ap.add_argument("--idle", type=float, default=0.0, help="seconds of GPU idle before the strict protocol")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    if a.idle > 0:
        torch.zeros(1 << 20, device="cuda").add_(1)
        torch.cuda.synchronize()
        time.sleep(a.idle)
        for t in range(a.trials):
            print(json.dumps(strict_trial(t)), flush=True)
        return
    if a.after_residency:
        sys.path.insert(0, str(ROOT))
        sys.path.insert(0, str(ROOT / "tests"))
        import test_verify_step_gpu as tv

        tv.test_fused_step_with_a_grid_larger_than_residency(0)
        print(json.dumps({"residency": "done"}), flush=True)
        for t in range(a.trials):
            print(json.dumps(strict_trial(t)), flush=True)
        return
    for t in range(a.trials):
        print(json.dumps(one_trial(t, dirty=t % 2 == 0)), flush=True)


if __name__ == "__main__":
    main()
