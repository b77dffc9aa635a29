"""Import shim for tests/test_reference_callers_gpu.py: a directory that provides ONLY the `kernels` package.

Put in front of the reference's ``src/`` on sys.path, ``import kernels`` resolves to this repository's drop-in
(``llm-inference-lab_b200/kernels``) while ``specdec``, ``scheduler`` ... stay the reference's own modules -- the
"kernel-only" integration of INTEGRATION.md, used to drive the reference's callers through these kernels."""
import pathlib as _pathlib

_real = _pathlib.Path(__file__).resolve().parents[3] / "llm-inference-lab_b200" / "kernels"
__path__ = [str(_real)]  # submodules (cuda_ops, registry, _capi, build) come from the real package directory
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
