#!/usr/bin/env python3
"""bench.py -- verify-step throughput of the B200-native speculative-decoding verify path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): verify-step tokens/s at B=64 per GPU, K=4, V=128256 (Llama-3.2 vocabulary),
bf16 logits, kv_append of the accepted rows for the 28 layers of Llama-3.2-3B (Hkv=8, D=128).
One step = fused verify over [B, K+1, V] logits + kv_append for all layers (+ the exchange of the packed
[B, K+2] records when N > 1) = ONE kernel launch (sd_verify_append), replayed from CUDA graphs.
tokens/s = (sum_b accept_len[b] + B) / t_step   (accepted drafts + one bonus token per sequence).

Timed region: all K steps are enqueued behind a device-side START GATE (sd_start_gate) that opens only when
(a) this rank's host has finished enqueuing them and (b) every rank has reached the gate; ev0 is recorded right
behind the gate, so neither host launch latency nor inter-rank start skew lands in a short (K = 20) window.

The JSON line follows the driver contract and adds `roofline` (dominant kernel: the fused step), `roofline_verify`,
`roofline_kv`, `roofline_attn`, `roofline_attn_d64`, `roofline_sample`, `cpu_baseline` and `e2e`.  See DESIGN.md.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import torch  # noqa: E402

METRIC = "verify-step tokens/s (B=64,K=4) at 1/2/4/8 B200; % of HBM roofline"
UNIT = "tokens/s"
SEED = 1234  # the reference's default seed (src/specdec/core/pipeline.py:407)
REF_SRC = ROOT / "baseline" / "_ref" / "src"  # unmodified copy of the reference (tools/make_baseline_ref.py)

# configs[2] of BASELINE.json: Llama-3.2-3B target (28 layers, 8 KV heads, head_dim 128, V=128256)
CFG = dict(B=64, K=4, V=128256, layers=28, Hq=24, Hkv=8, D=128, L=1024, p_match=0.8, nbuf=4)
STEP_KERNEL = "verify_step_kernel<__nv_bfloat16>"  # the instantiation sd_verify_append launches for bf16 logits


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the attention / sampler isolation loops (profiling runs)")
    ap.add_argument("--no-pipeline-baseline", action="store_true", help="skip the cfg-1 reference pipeline run (~25 s of CPU)")
    ap.add_argument("--tuning", default="", help="threads,vectors_per_thread for the verify kernel (sweeps)")
    ap.add_argument("--kv-variant", type=int, default=0)
    ap.add_argument("--batch", type=int, default=CFG["B"], help="sequences per GPU")
    ap.add_argument("--group", type=int, default=20, help="steps captured per CUDA graph")
    ap.add_argument("--no-gate", action="store_true", help="evidence runs: no start gate in front of the timed region")
    ap.add_argument("--no-fuse", action="store_true", help="two launches per step (verify, kv_append) instead of sd_verify_append")
    ap.add_argument("--no-zero-rejected", action="store_true",
                    help="evidence runs: rollback = length update only, rejected slots are not zeroed (not the reference's semantics)")
    ap.add_argument("--fused-mode", type=int, default=0, help="0 = lagged append (default), 1 = round-1 layout (append CTAs behind all verify CTAs)")
    ap.add_argument("--lag", type=int, default=0, help="fused mode 0: sequences of lag (0 = heuristic)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="two-launch form: kv_append waits for the whole verify grid (PDL) instead of per-sequence flags")
    ap.add_argument("--loopback-exchange", action="store_true",
                    help="evidence runs at N=1: the fused step with a one-rank peer exchange whose 'peer' is the GPU itself "
                         "(cost of the finalizer push + gather CTA without NVLink)")
    ap.add_argument("--exchange", default="peer", choices=["peer-kv", "peer", "nccl"],
                    help="N>1: peer-memory exchange, records pushed by the verify finalizers (peer, default) or by the "
                         "append part (peer-kv), or one NCCL all-gather per step (nccl)")
    return ap.parse_args()


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py): {short kernel name: entry}."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        return json.loads(p.read_text())
    except Exception:
        return {}


def traffic_for(traffic, short, must_contain):
    """The captured DRAM bytes of kernel `short`, but only if the capture is of the instantiation that is launched
    here (`must_contain` in its demangled name); otherwise None -- a figure from another build is not evidence."""
    e = traffic.get(short)
    if not e or must_contain not in e.get("kernel", ""):
        return None
    return e.get("dram_bytes")


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle reasons sampled IN PROCESS through NVML every ~1 ms while the GPU legs run (the timed
    region of a 20-step run lasts < 1 ms: an external `nvidia-smi -lms 20` never sees it).  Samples taken inside the
    timed region are counted separately."""

    def __init__(self, index: int):
        self.index, self.rows, self.thread, self.stop_flag = index, [], None, False
        self.in_region = False
        self.h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def sample(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:  # noqa: BLE001
            reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.rows.append((sm, reasons, self.in_region))

    def _run(self):
        while not self.stop_flag:
            try:
                self.sample()
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.001)

    def start(self):
        if self.h is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable: {getattr(self, 'err', '?')}"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(r[0] for r in self.rows)
        seen = set()
        for _, bits, _ in self.rows:
            for n, m in names.items():
                if bits & m:
                    seen.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(seen),
                "samples": len(sm), "samples_in_timed_region": sum(1 for r in self.rows if r[2]),
                "sm_mhz_min": sm[0] if sm else None, "how": "NVML in process, 1 ms period, while the GPU legs run"}


# ----------------------------------------------------------------------------------------- inputs
def make_inputs(B, K, V, p_match, seed, device):
    """Synthetic inputs per SURVEY.md section 8d: randn -> bf16 logits, randint ids, planted 10.0 with
    probability p_match (scripts/microbench_verify.py:57-64)."""
    g = torch.Generator(device=device).manual_seed(seed)
    logits = torch.randn(B, K + 1, V, generator=g, device=device, dtype=torch.float32).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, K), generator=g, device=device, dtype=torch.int64)
    plant = torch.rand(B, K, generator=g, device=device) < p_match
    bi, ki = torch.nonzero(plant, as_tuple=True)
    logits[bi, ki, ids[bi, ki]] = 10.0
    return logits, ids


def expected_accept(p, K):
    return sum(p ** i for i in range(1, K + 1))


# ----------------------------------------------------------------------------------------- reference arm
def reference_functions():
    """(verify_prefix_ref, kv_append_ref, kind): the UNMODIFIED reference functions from baseline/_ref when the copy
    is there (tools/make_baseline_ref.py), else the oracle's torch port (kind "port")."""
    if (REF_SRC / "kernels" / "reference.py").exists():
        import importlib.util

        spec = importlib.util.spec_from_file_location("_ref_kernels_reference", REF_SRC / "kernels" / "reference.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.verify_prefix_ref, mod.kv_append_ref, "reference"
    from oracle import torch_port as P  # (baseline/_ref missing: the pinned torch-CPU restatement, kind "port")

    return P.verify_prefix_port, P.kv_append_port, "port"


class ReferenceStep:
    """One verify step the way the reference's CPU path does it, all host threads:
    verify_prefix_ref on the K draft rows (torch.argmax + the per-sequence prefix loop, src/kernels/reference.py:13-56),
    the greedy bonus token of row accept_len[b] (src/specdec/core/pipeline.py:135-137: argmax), and kv_append_ref
    (torch.cat, reference.py:59-93) once per layer -- 28 calls, every one executed.  The 28 calls share ONE layer's base
    tensors (2 x 134 MB at B = 64, far beyond the last-level cache, so every call streams from DRAM like a distinct
    layer would; 28 distinct layers would need 7.5 GB of host memory and minutes of randn)."""

    def __init__(self, B, threads):
        torch.set_num_threads(threads)
        self.verify, self.kv_append, self.kind = reference_functions()
        K, V = CFG["K"], CFG["V"]
        g = torch.Generator().manual_seed(SEED)
        self.logits = torch.randn(B, K + 1, V, generator=g).to(torch.bfloat16)
        self.ids = torch.randint(0, V, (B, K), generator=g)
        plant = torch.rand(B, K, generator=g) < CFG["p_match"]
        bi, ki = torch.nonzero(plant, as_tuple=True)
        self.logits[bi, ki, self.ids[bi, ki]] = 10.0
        shp = (B, CFG["Hkv"], CFG["L"], CFG["D"])
        self.base_k = torch.randn(shp, generator=g).to(torch.bfloat16)
        self.base_v = torch.randn(shp, generator=g).to(torch.bfloat16)
        self.keep_rows = 4  # mean rows kept at p_match 0.8 is ~3.4 + 1; the reference appends a uniform slice
        self.new_k = torch.randn(B, CFG["Hkv"], self.keep_rows, CFG["D"], generator=g).to(torch.bfloat16)
        self.new_v = torch.randn(B, CFG["Hkv"], self.keep_rows, CFG["D"], generator=g).to(torch.bfloat16)
        self.B, self.K = B, K

    def step(self):
        t0 = time.perf_counter()
        acc, _ = self.verify(self.logits[:, : self.K], self.ids)
        rows = self.logits[torch.arange(self.B), acc.long()]
        bonus = torch.argmax(rows, dim=-1)
        t1 = time.perf_counter()
        for _ in range(CFG["layers"]):
            self.kv_append(self.base_k, self.base_v, self.new_k, self.new_v)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, int(acc.sum()) + self.B, bonus


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (baseline/_ref, unmodified), all host
    threads, rank 0 only.  Every one of the `--steps` steps (and `--warmup` warm-ups) is executed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    B = args.batch
    ref = ReferenceStep(B, cores)
    for _ in range(max(1, min(args.warmup, 3))):
        ref.step()
    tv = tk = 0.0
    toks = 0
    t_start = time.perf_counter()
    for _ in range(args.steps):
        a, b, n, _ = ref.step()
        tv, tk, toks = tv + a, tk + b, toks + n
    t_all = time.perf_counter() - t_start
    t_step = t_all / args.steps
    value = toks / t_all
    sample = (f"one host processes a {B}-sequence batch (CPU tokens/s does not depend on the batch: both legs are linear in B); "
              f"all {args.steps} steps executed: verify_prefix_ref on [B={B},K={CFG['K']},V={CFG['V']}] bf16 logits + argmax bonus "
              f"({tv / args.steps * 1e3:.1f} ms) + kv_append_ref (torch.cat, L={CFG['L']}, {ref.keep_rows} rows) x {CFG['layers']} layers "
              f"({tk / args.steps * 1e3:.1f} ms)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": ref.kind, "sample": sample,
                         "verify_ms": tv / args.steps * 1e3, "kv_append_ms": tk / args.steps * 1e3,
                         "cpu_model": cpu_model()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:  # noqa: BLE001
        pass
    return "?"


_PIPELINE_CODE = r'''
import os, sys, time, json, logging
os.environ["SPECDEC_ENABLE_KV_APPEND"] = "0"; os.environ["HF_HUB_OFFLINE"] = "1"
logging.disable(logging.CRITICAL)
sys.path.insert(0, sys.argv[1])
import torch
from transformers import GPT2Config, GPT2LMHeadModel
torch.set_num_threads(int(sys.argv[2]))
class Enc(dict):
    def to(s, d): return Enc({k: v.to(d) for k, v in s.items()})
    def __getattr__(s, k): return s[k]
class Tok:
    def __init__(s, v):
        s.vocab_size = v; s.eos_token_id = s.pad_token_id = s.bos_token_id = s.unk_token_id = v - 1
        s.pad_token = s.eos_token = "<eos>"; s.padding_side = "left"
    def __call__(s, text, **kw):
        texts = [text] if isinstance(text, str) else list(text)
        t = torch.tensor([[(hash((x, i)) % (s.vocab_size - 2)) for i in range(8)] for x in texts])
        return Enc(input_ids=t, attention_mask=torch.ones_like(t))
    def encode(s, text, return_tensors=None, **kw):
        r = s(text)["input_ids"]; return r if return_tensors else r[0].tolist()
    def decode(s, ids, **kw): return " ".join(str(int(i)) for i in (ids.tolist() if hasattr(ids, "tolist") else ids))
    def batch_decode(s, ids, **kw): return [s.decode(x) for x in ids]
    def __len__(s): return s.vocab_size
def tok(name, **kw): return Tok(50257)
def model(name, **kw):
    torch.manual_seed(1234)
    m = GPT2LMHeadModel(GPT2Config(n_layer=6) if "distil" in name else GPT2Config())
    dt = kw.get("torch_dtype") or kw.get("dtype")
    return m.to(dt) if dt is not None else m
import specdec.models.hf_wrappers as hw
hw.AutoTokenizer.from_pretrained = staticmethod(tok); hw.AutoModelForCausalLM.from_pretrained = staticmethod(model)
from specdec.core.pipeline import SpeculativePipeline
p = SpeculativePipeline(base_model="gpt2", draft_model="distilgpt2", max_draft=4, device="cpu", implementation="hf", seed=1234)
t0 = time.perf_counter()
out = p.generate("hello world", max_tokens=64, temperature=0.0, do_sample=False)
dt = time.perf_counter() - t0
print("RESULT " + json.dumps({"new_tokens": len(out["generated_tokens"]), "seconds": dt, "value": len(out["generated_tokens"]) / dt,
      "steps": out.get("steps"), "acceptance_rate": out.get("acceptance_rate"), "verify_ms": out.get("verification_time_ms"),
      "kv_append_enabled": out.get("kv_append_enabled")}))
'''


def reference_pipeline_baseline(threads):
    """BASELINE configs[0] / SURVEY 8(d) cfg-1: the reference's own SpeculativePipeline.generate (baseline/_ref, unmodified)
    with a GPT-2 target and a DistilGPT2 draft, random-init (no network: AutoModel/AutoTokenizer.from_pretrained are
    replaced by constructors from config and a stub tokenizer -- dependency injection at the transformers boundary), K = 4,
    64 new tokens, greedy, device cpu, SPECDEC_ENABLE_KV_APPEND=0 (the reference's KV path predates transformers 5)."""
    if not (REF_SRC / "specdec" / "core" / "pipeline.py").exists():
        return {"unavailable": "baseline/_ref missing (run tools/make_baseline_ref.py)"}
    import subprocess

    try:
        r = subprocess.run([sys.executable, "-c", _PIPELINE_CODE, str(REF_SRC), str(threads)], capture_output=True, text=True, timeout=240)
        for ln in r.stdout.splitlines():
            if ln.startswith("RESULT "):
                d = json.loads(ln[7:])
                d.update({"what": "reference SpeculativePipeline.generate, GPT-2 target + DistilGPT2 draft (random-init), K=4, 64 new tokens, "
                                  "greedy, device cpu (BASELINE configs[0])", "unit": "tokens/s", "cores": threads, "kind": "reference"})
                return d
        return {"unavailable": "reference pipeline did not finish: " + (r.stderr.strip().splitlines() or ["?"])[-1][:200]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"reference pipeline failed: {e}"[:240]}


def workload_config(args, world):
    return {
        "workload": "Llama-3.2-3B target verify step (BASELINE configs[2]): fused verify over [B,K+1,V] bf16 logits "
                    "+ kv_append of accepted rows, 28 layers x 8 KV heads x 128",
        "batch_per_gpu": args.batch, "global_batch": args.batch * world, "K": CFG["K"], "V": CFG["V"],
        "layers": CFG["layers"], "kv_heads": CFG["Hkv"], "head_dim": CFG["D"], "context_len": CFG["L"],
        "planted_acceptance": CFG["p_match"], "parallelism": f"batch-sharded dp{world}",
        "l2_policy": f"inputs rotate over {CFG['nbuf']} buffer sets ({CFG['nbuf']} x 82 MB logits + {CFG['nbuf']} x 37 MB KV rows "
                     "> 126 MB L2) so every replay reads from HBM",
        "launch": f"{args.group} steps per CUDA graph (PDL edges inside); timed region behind a device-side start gate; "
                  "2 untimed replays of the same graph after the requested warm-up steps"
                  + ("; with an exchange, 3 more untimed gated windows of one graph each" if world > 1 else ""),
    }


def bind_to_gpu_numa_node(dev):
    """Restrict this process to the CPUs of the GPU's NUMA node while the pinned staging buffers of the e2e leg are
    allocated and first touched (first-touch placement puts them on that node: host->device copies then do not cross
    the socket interconnect).  Returns (previous affinity, info dict) or (None, info) when the topology cannot be
    read or gives nothing to bind to -- nothing is changed in that case."""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = Path("/sys/bus/pci/devices") / bdf
        node = int((base / "numa_node").read_text().strip())
        cpus = set()
        for part in (base / "local_cpulist").read_text().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        prev = os.sched_getaffinity(0)
        want = cpus & prev
        if node < 0 or not want or want == prev:
            return None, {"node": node, "bound": False, "why": "single NUMA node / no narrower CPU set available"}
        os.sched_setaffinity(0, want)
        return prev, {"node": node, "bound": True, "cpus": len(want)}
    except Exception as e:  # noqa: BLE001  (no sysfs / no permission: leave the placement to the OS)
        return None, {"bound": False, "why": f"topology not readable: {e}"[:120]}


# ----------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch.distributed as dist

    from kernels import cuda_ops
    from scheduler.batch_shard import BatchShard, PeerExchange, StepResultExchange
    from specdec.cache.kv_arena import KVArena
    from specdec.verify_step import VerifyStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the sd_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.tuning:
        t, v = (int(x) for x in args.tuning.split(","))
        cuda_ops.set_verify_tuning(t, v)
    if args.kv_variant:
        cuda_ops.set_kv_variant(args.kv_variant)
    if args.fused_mode or args.lag:
        cuda_ops.set_fused_mode(args.fused_mode, args.lag)

    B, K, V, NB = args.batch, CFG["K"], CFG["V"], CFG["nbuf"]
    R = K + 1
    G = max(NB, (min(args.group, max(args.steps, NB)) // NB) * NB)  # steps per group graph, a multiple of NB
    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 20))
    # sequences are re-seated at context_len every RESEAT steps (a 256-byte device copy between graph
    # replays: "the batch finished, the next one starts"), so the arena stays small for any --steps
    RESEAT = max(G, (100 // G) * G)
    capacity = CFG["L"] + R * (RESEAT + 3 * G + 4 * NB + 16)
    arena = KVArena(CFG["layers"], B, CFG["Hkv"], CFG["D"], capacity, R, torch.bfloat16, dev)
    arena.set_lengths([CFG["L"]] * B)

    shard = BatchShard(B * world, world, rank)
    exch = peer = None
    exchange_kind = "none"
    if world == 1 and args.loopback_exchange:
        peer = PeerExchange.emulate(1, B, K, dev)[0]
        exchange_kind = "loopback (one-rank peer exchange on the GPU itself: evidence run)"
    if world > 1:
        if args.exchange in ("peer", "peer-kv"):
            try:
                peer = PeerExchange(shard, K, dev)
                exchange_kind = ("peer-memory stores issued by the append part once all records are final, gathered by the same CTA" if args.exchange == "peer-kv"
                                 else "peer-memory stores from the verify finalizers, confirmed + gathered by the closing CTA of the same launch")
            except Exception as e:  # noqa: BLE001  (CUDA IPC unavailable in this container)
                print(f"[rank {rank}] PeerExchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr, flush=True)
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer = None
            exch = StepResultExchange(shard, K, dev)
            exchange_kind = "ncclAllGather of the packed records"

    # NB rotating input sets; each set has its own VerifyStep (static buffers) sharing one arena
    steps, stage_k, stage_v = [], [], []
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    for i in range(NB):
        st = VerifyStep(B, K, V, arena=None, device=dev, exchange=peer)
        st.exchange_wait = False  # the append below (shared arena) confirms and gathers the records
        lg, ids = make_inputs(B, K, V, CFG["p_match"], SEED + 17 * rank + i, dev)
        st.logits.copy_(lg)
        st.draft_ids.copy_(ids)
        del lg
        steps.append(st)
        stage_k.append(torch.randn(arena.new_k.shape, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16))
        stage_v.append(torch.randn(arena.new_v.shape, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16))
    tables = [arena._build_table(stage_k[i], stage_v[i]) for i in range(NB)]

    def one_step(i, px=None):
        px = peer if px is None else px
        st = steps[i]
        arena.use_staging(stage_k[i], stage_v[i], tables[i])
        if not args.no_fuse:
            # the whole step in ONE launch (sd_verify_append)
            cuda_ops.verify_append_launch(st.logits, st.draft_ids, st.accept_len, st.accepted_mask, st.target_ids, st.bonus,
                                          st.packed, st.workspace, st.seq_ready, arena, zero_rejected=not args.no_zero_rejected, exchange=px,
                                          exchange_push_in_append=args.exchange == "peer-kv")
            if exch is not None:
                exch.all_gather(st.packed)
            return
        push_in_kv = args.exchange == "peer-kv"
        st.exchange = None if push_in_kv else px
        st.raise_flags = not args.no_overlap  # kv_append starts per sequence, overlapping the verify tail
        st.run()
        arena.append(st.accept_len, keep_bias=1, zero_rejected=True, check_capacity=False, exchange=px,
                     exchange_packed=st.packed if (push_in_kv and px is not None) else None,
                     seq_ready=st.seq_ready if st.raise_flags else None)
        if exch is not None:
            exch.all_gather(st.packed)

    # warm every buffer set eagerly, then capture (a) one graph per buffer set and (b) one GROUP graph holding G
    # consecutive steps (sets 0..NB-1 round robin): inside a graph the kernels are linked by programmatic
    # dependent-launch edges, so a kernel's launch latency overlaps the tail of its predecessor.
    for i in range(NB):
        one_step(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(dev)

    def capture(fn_list):
        with torch.cuda.stream(side):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                for fn in fn_list:
                    fn()
        return gr

    base_lens = torch.full((B,), CFG["L"], dtype=torch.int32, device=dev)

    state = {"since_reseat": 0}  # steps appended to the (shared) arena since its lengths were last re-seated

    def make_runner(fn_of_i, group=G, nsets=NB):
        """Returns run(n): n launches of fn_of_i over the rotating buffer sets."""
        singles = [capture([lambda i=i: fn_of_i(i)]) for i in range(nsets)]
        grp = capture([lambda i=i: fn_of_i(i % nsets) for i in range(group)])
        torch.cuda.synchronize()

        def run(n):
            k = n
            while k > 0:
                m = group if k >= group else 1
                if state["since_reseat"] + m > RESEAT:
                    arena.seq_lens.copy_(base_lens, non_blocking=True)
                    state["since_reseat"] = 0
                if m == group:
                    grp.replay()
                else:
                    singles[(n - k) % nsets].replay()
                state["since_reseat"] += m
                k -= m

        run.host_launches = lambda n: n // group + n % group + 2
        run.state = state
        return run

    run_steps = make_runner(one_step)
    arena.set_lengths([CFG["L"]] * B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gate = cuda_ops.StartGate(dev, peer)
    gate_box = {"g": gate}
    sampler = ClockSampler(local_rank)

    def timed(run, n):
        """barrier -> start gate (armed, held) -> ev0 -> n steps -> ev1, everything enqueued before the gate opens."""
        barrier()
        arena.seq_lens.copy_(base_lens, non_blocking=True)
        run.state["since_reseat"] = 0
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hold = (not args.no_gate) and run.host_launches(n) <= 64  # (a long queue behind a closed gate could fill the launch queue)
        if not args.no_gate:
            gate_box["g"].arm(hold=hold)
        sampler.in_region = True
        t0 = time.perf_counter()
        a.record()
        run(n)
        b.record()
        if hold:
            gate_box["g"].release()
        barrier()
        sampler.in_region = False
        return a.elapsed_time(b) / n, time.perf_counter() - t0

    sampler.start()
    run_steps(args.warmup)
    run_steps(2 * G)  # the graph of the timed region itself, untimed: its first replays pay one-off costs

    def warm_windows(run):
        """With an exchange, the first gated windows after a host synchronisation run slower than the steady state
        (one rank, loopback exchange, 20 steps: 29.6 / 28.9 / 28.7 / 28.5 / 28.5 us per step for consecutive windows): three
        untimed windows of the same protocol (barrier, gate, one group) go first."""
        if peer is not None:
            for _ in range(3):
                timed(run, G)

    warm_windows(run_steps)
    step_ms, t_wall = timed(run_steps, args.steps)
    if os.environ.get("SD_BENCH_DIAG") and peer is not None:  # evidence: is the first timed window special?
        t2, _ = timed(run_steps, args.steps)
        t3, _ = timed(run_steps, args.steps)
        gx = PeerExchange(shard, K, dev) if world > 1 else PeerExchange.emulate(1, B, K, dev)[0]
        gate_box["g"] = cuda_ops.StartGate(dev, gx)  # a gate that rendezvouses on its OWN buffer
        t4, _ = timed(run_steps, args.steps)
        t5, _ = timed(run_steps, args.steps)
        gate_box["g"] = gate
        t6, _ = timed(run_steps, args.steps)
        print(f"[diag rank {rank}] us/step: first {step_ms*1e3:.3f} again {t2*1e3:.3f} {t3*1e3:.3f} | gate on its own buffer "
              f"{t4*1e3:.3f} {t5*1e3:.3f} | shared again {t6*1e3:.3f}", file=sys.stderr, flush=True)

    # N>1: the same loop with the PIPELINED exchange (lag 1: a step gathers the previous step's records, so
    # the ranks are not forced into lockstep).  Reported next to the headline, which keeps lag 0.
    lag1_ms = 0.0
    if peer is not None:
        peer1 = PeerExchange(shard, K, dev, lag=1) if world > 1 else PeerExchange.emulate(1, B, K, dev, lag=1)[0]
        run_lag1 = make_runner(lambda i: one_step(i, peer1))
        run_lag1(args.warmup)
        run_lag1(2 * G)
        warm_windows(run_lag1)
        lag1_ms, _ = timed(run_lag1, args.steps)
        for st in steps:
            st.exchange = peer

    # ---- kernel-isolation loops for the roofline (same buffers, same rotation, CUDA events) ----
    def isolate(fn_of_i, n_launch, group=NB):
        run = make_runner(fn_of_i, group=group, nsets=group)
        run(max(2 * group, args.warmup))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(n_launch)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n_launch

    def kv_only(i):
        arena.use_staging(stage_k[i], stage_v[i], tables[i])
        arena.append(steps[i].accept_len, keep_bias=1, zero_rejected=True, check_capacity=False)

    n_iso = max(args.steps, 100)
    for st in steps:  # the kernel alone: without the peer stores / per-sequence flags (both need the append after it)
        st.exchange = None
        st.raise_flags = False
    verify_ms = isolate(lambda i: steps[i].run(), n_iso)
    for st in steps:
        st.exchange = peer
    arena.set_lengths([CFG["L"]] * B)
    kv_ms = isolate(kv_only, n_iso)

    extras = {}
    if not args.no_extras:
        lib = cuda_ops._capi.lib()
        # verify attention (kernel 3), one layer per launch, rotating over the arena's layers so that the
        # 2 x 134 MB of K/V per launch come from HBM; lengths pinned at L (the arena's own lengths moved)
        Hq = CFG["Hq"]
        attn_q = [torch.randn(B, Hq, R, CFG["D"], generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(NB)]
        attn_out = torch.empty_like(attn_q[0])
        attn_lens = torch.full((B,), CFG["L"], dtype=torch.int32, device=dev)
        attn_ws = torch.empty(int(lib.sd_verify_attention_workspace_bytes(B, Hq, R, CFG["D"], 0)), dtype=torch.uint8, device=dev)
        extras["attn_ms"] = isolate(lambda i: cuda_ops.verify_attention_cuda(
            attn_q[i], arena.k[i * 5 % CFG["layers"]], arena.v[i * 5 % CFG["layers"]], attn_lens, max_seq_len=CFG["L"],
            out=attn_out, workspace=attn_ws), n_iso)
        # balanced (persistent) form of the same kernel: one CTA per resident slot, equal spans of key tiles
        bal_ws = cuda_ops.BalancedAttentionWorkspace(B, Hq, CFG["Hkv"], R, CFG["D"], dev)
        extras["attn_bal_ms"] = isolate(lambda i: cuda_ops.verify_attention_cuda(
            attn_q[i], arena.k[i * 5 % CFG["layers"]], arena.v[i * 5 % CFG["layers"]], attn_lens, max_seq_len=CFG["L"],
            out=attn_out, balanced=bal_ws), n_iso)
        # the same kernel at head_dim 64 (Llama-3.2-1B: 32 query / 8 KV heads; the draft model of configs[2])
        D64, Hq64, Hkv64, L64 = 64, 32, 8, CFG["L"]
        k64 = [torch.randn(B, Hkv64, L64 + R, D64, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(8)]
        v64 = [torch.randn(B, Hkv64, L64 + R, D64, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(8)]
        q64 = [torch.randn(B, Hq64, R, D64, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16) for _ in range(NB)]
        out64 = torch.empty_like(q64[0])
        ws64 = torch.empty(int(lib.sd_verify_attention_workspace_bytes(B, Hq64, R, D64, 0)), dtype=torch.uint8, device=dev)
        # (8 K/V sets = 8 x 2 x 67 MB > L2; graph of 8 launches: launch j reads set j)
        extras["attn64_ms"] = isolate(lambda j: cuda_ops.verify_attention_cuda(
            q64[j % NB], k64[j], v64[j], attn_lens, max_seq_len=L64, out=out64, workspace=ws64), (n_iso // 8) * 8, group=8)
        bal_ws64 = cuda_ops.BalancedAttentionWorkspace(B, Hq64, Hkv64, R, D64, dev)
        extras["attn64_bal_ms"] = isolate(lambda j: cuda_ops.verify_attention_cuda(
            q64[j % NB], k64[j], v64[j], attn_lens, max_seq_len=L64, out=out64, balanced=bal_ws64), (n_iso // 8) * 8, group=8)
        extras["attn64_bytes"] = 2 * B * Hkv64 * (L64 + R) * D64 * 2 + 2 * B * Hq64 * R * D64 * 2
        del k64, v64
        # bonus sampler (kernel 4) on row accept_len[b] of the same logits, reference defaults T=0.7, k=50, p=0.9
        uni = torch.rand(B, device=dev)
        samp_out = torch.empty(B, dtype=torch.int32, device=dev)
        extras["sample_ms"] = isolate(lambda i: cuda_ops.sample_bonus_cuda(
            steps[i].logits, row_index=steps[i].accept_len, uniform=uni, temperature=0.7, top_k=50, top_p=0.9, out=samp_out), n_iso)
    clocks = sampler.stop()

    acc_sum = [int(steps[i].accept_len.sum().item()) for i in range(NB)]
    tokens_per_step = sum(acc_sum[s % NB] + B for s in range(args.steps)) / args.steps
    mean_keep = sum(a + B for a in acc_sum) / (NB * B)
    assert int(arena.overflow.item()) == 0, "KV arena overflowed during the bench"
    dev_err = cuda_ops.device_error()
    assert dev_err == 0, f"a bounded device-side wait expired during the bench (sites {dev_err})"

    # ---- e2e: host buffers in, packed record out, copies inside the timed region ----
    e2e = None
    if e2e_steps:
        arena.set_lengths([CFG["L"]] * B)
        e2e_graph = capture([lambda: one_step(0)])
        prev_affinity, numa_info = bind_to_gpu_numa_node(dev)
        h_logits = [steps[i].logits.cpu().pin_memory() for i in range(2)]
        h_ids = [steps[i].draft_ids.cpu().pin_memory() for i in range(2)]
        h_k = [stage_k[i].cpu().pin_memory() for i in range(2)]
        h_v = [stage_v[i].cpu().pin_memory() for i in range(2)]
        h_out = torch.empty((B, K + 2), dtype=torch.int32).pin_memory()
        h2d = h_logits[0].numel() * 2 + h_ids[0].numel() * 8 + 2 * h_k[0].numel() * 2
        d2h = h_out.numel() * 4

        # the 118.8 MB of inputs go up on two copy streams (logits halves + K rows | V rows + ids): one stream
        # alone leaves PCIe bandwidth unused between its back-to-back copies
        cp1, cp2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        half = B // 2

        def e2e_step(s):
            i = s % 2
            cur = torch.cuda.current_stream()
            cp1.wait_stream(cur)
            cp2.wait_stream(cur)
            with torch.cuda.stream(cp1):
                steps[0].logits[:half].copy_(h_logits[i][:half], non_blocking=True)
                stage_k[0].copy_(h_k[i], non_blocking=True)
            with torch.cuda.stream(cp2):
                steps[0].logits[half:].copy_(h_logits[i][half:], non_blocking=True)
                stage_v[0].copy_(h_v[i], non_blocking=True)
                steps[0].draft_ids.copy_(h_ids[i], non_blocking=True)
            cur.wait_stream(cp1)
            cur.wait_stream(cp2)
            e2e_graph.replay()
            h_out.copy_(steps[0].packed, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller needs the record to advance sequences
            return int(h_out[:, 0].sum()) + B

        for s in range(3):
            e2e_step(s)
        barrier()
        t0 = time.perf_counter()
        toks = 0
        for s in range(e2e_steps):
            toks += e2e_step(s)
        barrier()
        t_e2e = time.perf_counter() - t0
        e2e = {"tokens": toks, "seconds": t_e2e, "h2d": h2d, "d2h": d2h, "numa": numa_info}
        if prev_affinity is not None:
            try:
                os.sched_setaffinity(0, prev_affinity)  # the cpu_baseline leg uses every core
            except Exception:  # noqa: BLE001
                pass

    # ---- reduce over ranks: max time, summed tokens ----
    attn_ms = extras.get("attn_ms", 0.0)
    if world > 1:
        t = torch.tensor([step_ms, verify_ms, kv_ms, attn_ms, e2e["seconds"] if e2e else 0.0, lag1_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, verify_ms, kv_ms, attn_ms, e2e_s, lag1_ms = (float(x) for x in t.tolist())
        c = torch.tensor([tokens_per_step, float(e2e["tokens"]) if e2e else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tokens_per_step_all, e2e_tokens_all = (float(x) for x in c.tolist())
    else:
        tokens_per_step_all = tokens_per_step
        e2e_s = e2e["seconds"] if e2e else 0.0
        e2e_tokens_all = float(e2e["tokens"]) if e2e else 0.0

    if rank == 0:
        peak, peak_src = load_peaks()
        verify_bytes = B * R * V * 2
        kv_bytes = arena.step_bytes(mean_keep, R, not args.no_zero_rejected)
        v_gbs = verify_bytes / (verify_ms * 1e-3) / 1e9
        kv_gbs = kv_bytes / (kv_ms * 1e-3) / 1e9
        traffic = load_traffic()
        line = {
            "metric": METRIC, "value": tokens_per_step_all / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "tokens_per_step": tokens_per_step_all, "positions_per_s": B * world * R / (step_ms * 1e-3),
            "mean_accept_len": mean_keep - 1.0,
            "gpu_launches": args.steps * (2 if args.no_fuse else 1),
            "launches_per_step": ({"verify_bulk_kernel": 1, "kv_append_col_kernel": 1} if args.no_fuse else {STEP_KERNEL: 1}) | {
                                  "nccl_allgather": 1 if exch is not None else 0},
            "start_gate": "off" if args.no_gate else "sd_start_gate (host 'all enqueued' flag + peer arrival words), ev0 recorded behind it",
            "exchange": exchange_kind,
            "pipelined_exchange": ({"lag_steps": 1, "ms_per_step": lag1_ms, "value": tokens_per_step_all / (lag1_ms * 1e-3),
                                    "note": "same step, but each launch gathers the PREVIOUS step's records: ranks are not in lockstep; "
                                            "every rank still has its own shard's records immediately"} if lag1_ms > 0 else None),
            "roofline": None,
            "roofline_verify": {"bound": "hbm", "kernel": "verify_bulk_kernel<__nv_bfloat16, false> (one cp.async.bulk per CTA chunk), timed alone", "achieved": v_gbs, "peak": peak, "unit": "GB/s",
                                "frac": v_gbs / peak, "frac_of_8TBs_nominal": v_gbs / 8000.0,
                                "traffic": traffic_for(traffic, "verify_bulk_kernel", "verify_bulk_kernel<__nv_bfloat16"),
                                "algorithmic_bytes": verify_bytes, "us_per_launch": verify_ms * 1e3, "peak_source": peak_src},
            "roofline_kv": {"bound": "hbm", "kernel": "kv_append_col_kernel<5> (all 28 layers, one launch), timed alone", "achieved": kv_gbs, "peak": peak, "unit": "GB/s",
                            "frac": kv_gbs / peak, "algorithmic_bytes": kv_bytes, "us_per_launch": kv_ms * 1e3,
                            "mean_rows_kept": mean_keep, "traffic": traffic_for(traffic, "kv_append_col_kernel", "kv_append_col_kernel<5>")},
            "step_hbm_frac": (verify_bytes + kv_bytes) / (step_ms * 1e-3) / 1e9 / peak,
            "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        if "attn_ms" in extras:
            attn_bytes = 2 * B * CFG["Hkv"] * (CFG["L"] + R) * CFG["D"] * 2 + 2 * B * CFG["Hq"] * R * CFG["D"] * 2
            attn_flops = 4 * B * CFG["Hq"] * R * (CFG["L"] + R) * CFG["D"]
            # both forms of the kernel are timed (rank 0's own numbers); the line quotes the faster one and names it
            forms = {"standard (one CTA per (sequence, KV head))": attn_ms, "balanced (persistent, equal spans of key tiles)": extras["attn_bal_ms"]}
            best = min(forms, key=forms.get)
            attn_best = forms[best]
            attn_gbs = attn_bytes / (attn_best * 1e-3) / 1e9
            line["roofline_attn"] = {"bound": "hbm", "kernel": "verify_attn_kernel<bf16,D=128> (one layer per launch, mma.sync m16n8k16)",
                                     "form": best, "us_per_launch_by_form": {k: v * 1e3 for k, v in forms.items()},
                                     "achieved": attn_gbs, "peak": peak, "unit": "GB/s", "frac": attn_gbs / peak,
                                     "algorithmic_bytes": attn_bytes, "us_per_launch": attn_best * 1e3,
                                     "tflops": attn_flops / (attn_best * 1e-3) / 1e12,
                                     "traffic": traffic_for(traffic, "verify_attn_kernel", "verify_attn_kernel<__nv_bfloat16, 128") if best.startswith("standard") else None,
                                     "config": {"B": B, "Hq": CFG["Hq"], "Hkv": CFG["Hkv"], "R": R, "D": CFG["D"], "context_len": CFG["L"]}}
            forms64 = {"standard (one CTA per (sequence, KV head))": extras["attn64_ms"], "balanced (persistent, equal spans of key tiles)": extras["attn64_bal_ms"]}
            best64 = min(forms64, key=forms64.get)
            g64 = extras["attn64_bytes"] / (forms64[best64] * 1e-3) / 1e9
            line["roofline_attn_d64"] = {"bound": "hbm", "kernel": "verify_attn_kernel<bf16,D=64> (Llama-3.2-1B geometry: 32 query / 8 KV heads)",
                                         "form": best64, "us_per_launch_by_form": {k: v * 1e3 for k, v in forms64.items()},
                                         "achieved": g64, "peak": peak, "unit": "GB/s", "frac": g64 / peak,
                                         "algorithmic_bytes": extras["attn64_bytes"], "us_per_launch": forms64[best64] * 1e3,
                                         "config": {"B": B, "Hq": 32, "Hkv": 8, "R": R, "D": 64, "context_len": CFG["L"]}}
            samp_bytes = B * V * 2
            sg = samp_bytes / (extras["sample_ms"] * 1e-3) / 1e9
            line["roofline_sample"] = {"bound": "hbm", "kernel": "sample_bonus kernel (temperature 0.7, top-k 50, top-p 0.9; one logits row per sequence)",
                                       "achieved": sg, "peak": peak, "unit": "GB/s", "frac": sg / peak, "algorithmic_bytes": samp_bytes,
                                       "us_per_launch": extras["sample_ms"] * 1e3,
                                       "note": "algorithmic bytes = one read of the B selected rows"}
        if args.no_fuse:
            line["roofline"] = dict(line["roofline_verify"], kernel="verify_bulk_kernel<bf16> (dominant kernel of the two-launch step)")
        else:
            # the timed region launches ONE kernel per step: the fused verify + append kernel is the dominant kernel
            step_bytes = verify_bytes + kv_bytes
            step_gbs = step_bytes / (step_ms * 1e-3) / 1e9
            kname = STEP_KERNEL if args.fused_mode == 0 else "verify_append_kernel<__nv_bfloat16, 3>"
            line["roofline"] = {"bound": "hbm", "kernel": kname + " (the whole step in one launch: fused verify + kv_append of 28 layers)",
                                "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                                "frac_of_8TBs_nominal": step_gbs / 8000.0,
                                "traffic": traffic_for(traffic, kname.split("<")[0], kname),
                                "algorithmic_bytes": step_bytes, "us_per_launch": step_ms * 1e3, "peak_source": peak_src,
                                "note": "algorithmic bytes = logits read once (82.08 MB) + accepted KV rows read + offered slots written "
                                        f"({kv_bytes / 1e6:.1f} MB at {mean_keep:.2f} rows kept); duration = CUDA-event time of the timed region / launches; "
                                        "traffic is null unless profiles/ncu_traffic.json holds a capture of exactly this instantiation"}
        if e2e:
            line["e2e"] = {"value": e2e_tokens_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                           "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                           "api": "VerifyStep static buffers <- pinned host tensors; packed [B,K+2] record -> host",
                           "host_numa": e2e.get("numa")}
        if not args.no_cpu_baseline and world == 1:  # reported baseline: rank 0 at N=1 only
            cores = os.cpu_count() or 1
            ref = ReferenceStep(B, cores)
            ref.step()
            reps, tv, tk, toks = 0, 0.0, 0.0, 0
            t0 = time.perf_counter()
            while reps < 12 and (reps < 3 or time.perf_counter() - t0 < 8.0):
                a_, b_, n_, _ = ref.step()
                tv, tk, toks, reps = tv + a_, tk + b_, toks + n_, reps + 1
            line["cpu_baseline"] = {
                "value": toks / (tv + tk), "unit": UNIT, "cores": cores, "kind": ref.kind, "cpu_model": cpu_model(),
                "sample": (f"{reps} whole steps on rank 0's host, all executed: verify_prefix_ref on [{B},{K},{V}] bf16 logits + argmax bonus "
                           f"({tv / reps * 1e3:.1f} ms) + kv_append_ref (torch.cat, L={CFG['L']}, {ref.keep_rows} rows) x {CFG['layers']} layer calls "
                           f"({tk / reps * 1e3:.1f} ms; the calls share one layer's tensors, 2 x {B * CFG['Hkv'] * CFG['L'] * CFG['D'] * 2 / 1e6:.0f} MB, far beyond the LLC)"),
            }
            del ref
            if not args.no_pipeline_baseline:
                line["cpu_baseline"]["e2e_pipeline"] = reference_pipeline_baseline(cores)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: destroying a communicator whose collectives are still held by
        # live CUDA graphs has been seen to hang (round 1, N=2: the JSON line was out, the exit never came).
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
