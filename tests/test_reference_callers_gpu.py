"""GPU: the REFERENCE's own callers (unmodified, baseline/_ref/src) driving this package's kernels.

Each test runs in a subprocess whose sys.path is [tests/_kernels_only, baseline/_ref/src]: ``kernels`` is this
repository's drop-in, every other module (``specdec.policies``, ``specdec.models.hf_wrappers``,
``specdec.core.pipeline`` ...) is the reference's.  Compared with the same callers forced onto their PyTorch path
(``SPECDEC_FORCE_PYTORCH_BACKEND=1`` / ``_append_kv_pytorch``): accept lengths, KV tensors and generated tokens
must be identical, and the kernel path must really have been taken (``verify_backend == "cuda"``, arena-tagged
results).  baseline/_ref is created by tools/make_baseline_ref.py (and __graft_entry__.build()) in the build
container and travels with the repo snapshot; /root/reference itself is never read here.

reference: src/specdec/policies/policies.py:101-204 (accept_tokens), src/specdec/models/hf_wrappers.py:914-1030
(append_kv_cache / _append_kv_with_kernel), src/specdec/core/pipeline.py:893 (generate)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
REF_SRC = ROOT / "baseline" / "_ref" / "src"
SHIM = ROOT / "tests" / "_kernels_only"

needs_ref = pytest.mark.skipif(not (REF_SRC / "specdec" / "policies" / "policies.py").exists(),
                               reason="baseline/_ref missing: run tools/make_baseline_ref.py in the build container")

PRELUDE = r'''
import os, sys, json, logging
logging.disable(logging.WARNING)
OURS, REF = sys.argv[1], sys.argv[2]
import torch
import kernels
assert kernels.__path__[0].startswith(OURS), kernels.__path__
info = kernels.get_kernel_info()
assert info["verify_backend"] == "cuda" and info["kv_append_backend"] == "cuda", info
def from_ref(m): assert m.__file__.startswith(REF), m.__file__
'''


def _run(body: str, timeout=900):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(SHIM), str(REF_SRC)])
    env["HF_HUB_OFFLINE"] = "1"
    env.pop("SPECDEC_FORCE_PYTORCH_BACKEND", None)
    r = subprocess.run([sys.executable, "-c", PRELUDE + body, str(ROOT / "llm-inference-lab_b200"), str(REF_SRC)],
                       capture_output=True, text=True, env=env, cwd="/tmp", timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-5000:]
    last = r.stdout.strip().splitlines()[-1]
    assert last.startswith("ok"), r.stdout[-2000:]
    return last


@needs_ref
def test_reference_longest_prefix_policy_uses_these_kernels():
    out = _run(r'''
import specdec.policies.policies as pol
from_ref(pol)
policy = pol.LongestPrefixPolicy()
assert policy.kernels_available and policy.kernel_info["verify_backend"] == "cuda"
g = torch.Generator().manual_seed(0)
n_cases = 0
for case in range(48):
    K = 1 + case % 8
    V = [50257, 128256, 1000, 32000][case % 4]
    dtype = [torch.float32, torch.bfloat16, torch.float16][case % 3]
    B = 1 + case % 3
    logits_all = torch.randn(B, K, V, generator=g).to(dtype).cuda()
    ids_all = torch.randint(0, V, (B, K), generator=g).cuda()
    plant = torch.rand(B, K, generator=g) < 0.75
    for b in range(B):
        for k in range(K):
            if plant[b, k]:
                logits_all[b, k, ids_all[b, k]] = 10.0
    for b in range(B):  # the pipeline calls the policy once per sequence on a batch slice (pipeline.py:2960-3018)
        logits, ids = logits_all[b:b + 1], ids_all[b:b + 1]
        base_tokens = torch.argmax(logits, -1)
        os.environ["SPECDEC_FORCE_PYTORCH_BACKEND"] = "0"
        n1, i1 = policy.accept_tokens(ids, base_tokens, None, logits)
        os.environ["SPECDEC_FORCE_PYTORCH_BACKEND"] = "1"
        n2, i2 = policy.accept_tokens(ids, base_tokens, None, logits)
        assert i1["verify_backend"] == "cuda", i1   # the kernel path was taken (no silent except-fallback)
        assert i2["verify_backend"] == "torch", i2
        assert n1 == n2 == i1["accepted_len"], (case, b, n1, n2)
        n_cases += 1
print("ok", n_cases)
''')
    assert int(out.split()[1]) >= 90


@needs_ref
def test_reference_hfwrapper_kv_append_uses_these_kernels():
    _run(r'''
import specdec.models.hf_wrappers as hw
import specdec.cache.kv_types as kvt
from_ref(hw); from_ref(kvt)
class _Model:
    config = object()
    def forward(self): pass
def wrapper():
    w = object.__new__(hw.HFWrapper)   # (the constructor downloads weights; the KV path needs none)
    w.logger = logging.getLogger("t"); w._device = "cuda"; w._kv_append_enabled = True
    w._model = _Model(); w._kv_cache = None; w._last_generated_kv = None
    return w
def cache(layers, n, seed, B=3, H=4, D=64, dtype=torch.float16):
    g = torch.Generator().manual_seed(seed)
    kv = tuple((torch.randn(B, H, n, D, generator=g).to(dtype).cuda(), torch.randn(B, H, n, D, generator=g).to(dtype).cuda())
               for _ in range(layers))
    return kvt.KVCache(past_key_values=kv, seq_len=n, dtype=dtype, device=torch.device("cuda"))
for dtype in (torch.float16, torch.bfloat16, torch.float32):
    w, ref = wrapper(), wrapper()
    assert w.supports_kv_append()
    sizes = [9, 4, 1, 5, 3, 8, 2]
    for i, n in enumerate(sizes):
        c = cache(3, n, 10 + i, dtype=dtype)
        w.append_kv_cache(c)                                   # get_kv_append("cuda") -> this repo's kv_append_cuda
        ref._kv_cache = c if ref._kv_cache is None else ref._append_kv_pytorch(ref._kv_cache, c)   # torch.cat
    got, want = w.get_kv_cache(), ref.get_kv_cache()
    assert got.seq_len == want.seq_len == sum(sizes)
    for (gk, gv), (wk, wv) in zip(got.past_key_values, want.past_key_values):
        assert torch.equal(gk, wk) and torch.equal(gv, wv)
        info = kernels.cuda_ops.arena_info(gk)
        assert info is not None and info[0] == sum(sizes), "result did not come from the sd_b200 kv_append (fallback taken?)"
print("ok")
''')


@needs_ref
def test_reference_pipeline_generates_the_same_tokens_with_these_kernels():
    _run(r'''
import specdec.core.pipeline as pl
import specdec.policies.policies as pol
from_ref(pl); from_ref(pol)
def gen(force_torch):
    os.environ["SPECDEC_FORCE_PYTORCH_BACKEND"] = "1" if force_torch else "0"
    p = pl.SpeculativePipeline(implementation="fake", device="cuda", max_draft=4, seed=1234)
    torch.manual_seed(1234)
    out = p.generate("the quick brown fox", max_tokens=48, temperature=0.0, do_sample=False)
    return out, p
a, pa = gen(False)
b, pb = gen(True)
assert pa.policy.kernel_info["verify_backend"] == "cuda", pa.policy.kernel_info
assert a["generated_tokens"] == b["generated_tokens"], (a["generated_tokens"], b["generated_tokens"])
assert a["accepted"] == b["accepted"] and a["proposed"] == b["proposed"]
assert len(a["generated_tokens"]) > 0
print("ok", len(a["generated_tokens"]), a["accepted"], a["proposed"])
''')
