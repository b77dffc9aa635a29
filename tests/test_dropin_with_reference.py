"""The zero-change drop-in of INTEGRATION.md section 1, checked against the real reference tree when it is
available (this container; skipped on the GPU box, where /root/reference does not exist): with this repo's
package directory in front of the reference's ``src/`` on ``sys.path``, ``kernels`` and the mirrored modules
resolve here, everything else (pipeline, models, controllers, token validation ...) to the reference."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("SD_REFERENCE_ROOT", "/root/reference"))

PROBE = r'''
import sys
ours, ref = sys.argv[1], sys.argv[2]
import kernels, scheduler, specdec
import specdec.core.pipeline as pipeline
import specdec.policies.policies as pol
import specdec.policies.controllers as ctl
import specdec.cache as cache
import specdec.cache.kv_types as kvt
import specdec.utils as utils
import specdec.utils.graph_capture as gc
import specdec.utils.token_validation as tv
from specdec.models.hf_wrappers import HFWrapper
def here(m): return m.__file__.startswith(ours)
def there(m): return m.__file__.startswith(ref)
assert here(kernels) and here(pol) and here(kvt) and here(gc) and here(scheduler) and here(tv), "mirrored modules must resolve to this repo"
assert there(pipeline) and there(ctl), "everything else must resolve to the reference"
assert pipeline.SafeKVCacheManager is cache.SafeKVCacheManager and cache.SafeKVCacheManager.__module__ == "specdec.cache.kv_cache_manager"
assert pipeline.create_policy is pol.create_policy and pipeline.AcceptancePolicy is pol.AcceptancePolicy
assert specdec.SpeculativePipeline is pipeline.SpeculativePipeline
assert utils.validate_and_clamp_tokens is tv.validate_and_clamp_tokens
from specdec.policies import KController, create_controller, create_policy
assert KController is ctl.KController
assert scheduler.SpeculativeScheduler.__module__ == "scheduler.speculative_scheduler"
info = kernels.get_kernel_info()
assert info["arch"] == "sm_100a" and "library" in info
print("ok", info["verify_backend"])
'''


@pytest.mark.skipif(not (REF / "src" / "specdec" / "core" / "pipeline.py").exists(), reason="reference tree not present")
def test_package_dir_in_front_of_reference_src():
    env = dict(os.environ)
    ours = str(ROOT / "llm-inference-lab_b200")
    env["PYTHONPATH"] = os.pathsep.join([ours, str(REF / "src")])
    r = subprocess.run([sys.executable, "-c", PROBE, ours, str(REF / "src")], capture_output=True, text=True, env=env, cwd="/tmp", timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.strip().splitlines()[-1].startswith("ok")
