"""Build-time evidence read from the SASS of the in-tree library (no GPU needed; needs cuobjdump).

ptxas decides on its own whether the VPT independent 128-bit loads of the verify kernel's batch stay
ahead of their first use; once it serialised load -> use (one request in flight per thread) after a
harmless refactor and the kernel lost 20 % of its bandwidth.  The launch bounds in sd_verify.cu pin the
schedule; this test keeps it pinned."""
import re
import shutil
import subprocess

import pytest

from kernels import build

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _sass(fun: str) -> str:
    lib = build.build_cuda_kernels()
    r = subprocess.run([CUOBJDUMP, "-sass", "-fun", fun, str(lib)], capture_output=True, text=True)
    if r.returncode != 0 or "Function" not in r.stdout:
        pytest.skip(f"cuobjdump could not list {fun}")
    return r.stdout


@pytest.mark.skipif(not shutil.which(CUOBJDUMP), reason="cuobjdump not installed")
def test_default_verify_kernel_fetches_its_chunk_with_one_bulk_copy():
    """The default variant does not depend on ptxas keeping a load batch together: the chunk arrives by a
    single UBLKCP and is consumed from shared memory (LDS.128) after an mbarrier wait (SYNCS)."""
    sass = _sass("_ZN2sd18verify_bulk_kernelI13__nv_bfloat16Lb0EEEvNS_12VerifyParamsE")
    assert len(re.findall(r"UBLKCP", sass)) == 1
    assert "SYNCS" in sass and "LDS.128" in sass and "HMNMX2.BF16_V2.NAN" in sass
    assert "LDG.E.NA.128" not in sass  # no register-staged streaming loads on this path


@pytest.mark.skipif(not shutil.which(CUOBJDUMP), reason="cuobjdump not installed")
@pytest.mark.parametrize("vpt", [4, 8])
def test_register_staged_variant_issues_several_loads_before_first_use(vpt):
    """Variant 0 (kept for evidence runs): at least two 128-bit loads in flight before the first use.
    (ptxas has been seen to serialise further when register pressure changes -- the reason the default
    moved to the bulk-copy variant.)"""
    sass = _sass(f"_ZN2sd13verify_kernelI13__nv_bfloat16Li{vpt}ELb0EEEvNS_12VerifyParamsE")
    ops = re.findall(r"(LDG\.E\.NA\.128|HMNMX2)", sass)
    assert ops.count("LDG.E.NA.128") == vpt
    assert ops.index("HMNMX2") >= 2, "".join("L" if o.startswith("LDG") else "m" for o in ops)


@pytest.mark.skipif(not shutil.which(CUOBJDUMP), reason="cuobjdump not installed")
def test_blackwell_native_instructions_present():
    """TMA bulk copies (UBLKCP), mbarrier waits (SYNCS), PDL (ACQBULK/PREEXIT), tensor-core MMA (HMMA)
    and ldmatrix (LDSM) must be in the library -- the evidence table of B200_PROFILING.md."""
    lib = build.build_cuda_kernels()
    r = subprocess.run([CUOBJDUMP, "-sass", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0
    for mnemonic in ("UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "HMMA.16816.F32.BF16", "LDSM", "LDGSTS"):
        assert mnemonic in r.stdout, f"{mnemonic} not found in the SASS of {lib}"


@pytest.mark.skipif(not shutil.which(CUOBJDUMP), reason="cuobjdump not installed")
@pytest.mark.parametrize("kernel", ["verify_step_kernelI13__nv_bfloat16EE", "verify_append_kernelI13__nv_bfloat16Li3EE"],
                         ids=["lagged_append_default", "round1_layout"])
def test_fused_step_kernel_keeps_four_ctas_per_sm(kernel):
    """The shipped fused step kernel (verify_step_kernel<bf16>: lagged append, rows parked in shared memory) and the
    round-1 layout kept for A/B must stay at <= 64 registers without a stack frame: 256 threads x 64 registers = 4
    resident CTAs per SM (an 80-register build of round 1 held 3 and measured 33.5 us per step against 32.6 us; three
    CTAs per SM measured 33.2 us against 31.1 us in round 2)."""
    lib = build.build_cuda_kernels()
    r = subprocess.run([CUOBJDUMP, "--dump-resource-usage", str(lib)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump could not dump resource usage")
    lines = r.stdout.splitlines()
    hits = [lines[i + 1] for i, ln in enumerate(lines) if kernel in ln and i + 1 < len(lines)]
    assert hits, f"{kernel} not found in the library"
    regs = int(re.search(r"REG:(\d+)", hits[0]).group(1))
    stack = int(re.search(r"STACK:(\d+)", hits[0]).group(1))
    # The shipped kernel must have no stack frame at all.  The A/B kernel may spill ONE register: since the exchange
    # gather polls 12 words per thread, ptxas keeps one 4-byte value of its (once-per-launch) gather branch on the stack;
    # residency (registers, shared memory) is what the comparison needs and that is unchanged.
    assert regs <= 64 and stack <= (0 if "verify_step_kernel" in kernel else 8), hits[0]


@pytest.mark.skipif(not shutil.which(CUOBJDUMP), reason="cuobjdump not installed")
def test_fused_step_kernel_parks_rows_with_cp_async():
    """verify_step_kernel<bf16>: one UBLKCP for the logits chunk, LDGSTS (cp.async) for the carried KV rows, packed
    NaN-propagating max for the fold."""
    sass = _sass("_ZN2sd18verify_step_kernelI13__nv_bfloat16EEvNS_12VerifyParamsENS_8KvParamsEiii")
    assert len(re.findall(r"UBLKCP", sass)) == 1
    assert "LDGSTS" in sass and "HMNMX2.BF16_V2.NAN" in sass
