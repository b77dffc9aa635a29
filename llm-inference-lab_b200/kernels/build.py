"""Ahead-of-time build of the sm_100a kernel library.

Mirrors the role of the reference's ``src/kernels/build.py`` (JIT via
``torch.utils.cpp_extension.load`` with ``-arch=sm_{major}{minor}``, build.py:88-111, and a
source-hash cache key, build.py:35-38) with a B200-only recipe: every ``csrc/*.cu`` is compiled by
``nvcc -gencode arch=compute_100a,code=sm_100a`` into ONE in-tree shared library
(``llm-inference-lab_b200/lib/libsd_verify_b200.so``) that exports the C ABI of
``include/sd_verify.h``.  There is no Triton / PyTorch fallback chain: if the library cannot be
built or loaded, ``load_kernels`` raises.

Environment:
  SD_B200_NVCC        path to nvcc (default: $CUDA_HOME/bin/nvcc, /usr/local/cuda/bin/nvcc, PATH)
  SD_B200_REBUILD=1   ignore the hash cache
"""

from __future__ import annotations

import hashlib
import logging
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional

logger = logging.getLogger(__name__)

PKG_ROOT = Path(__file__).resolve().parent.parent
REPO_ROOT = PKG_ROOT.parent
CSRC = PKG_ROOT / "csrc"
LIB_DIR = PKG_ROOT / "lib"
LIB_PATH = LIB_DIR / "libsd_verify_b200.so"
HASH_PATH = LIB_DIR / "libsd_verify_b200.hash"
INCLUDE = REPO_ROOT / "include"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def get_cuda_arch() -> str:
    """The only architecture this library is written for (reference: build.py:19-32 queried the device)."""
    return "sm_100a"


def find_nvcc() -> Optional[str]:
    cands = [os.environ.get("SD_B200_NVCC")]
    if os.environ.get("CUDA_HOME"):
        cands.append(os.path.join(os.environ["CUDA_HOME"], "bin", "nvcc"))
    cands += ["/usr/local/cuda/bin/nvcc", shutil.which("nvcc")]
    for c in cands:
        if c and os.path.isfile(c) and os.access(c, os.X_OK):
            return c
    return None


def sources() -> List[Path]:
    return sorted(CSRC.glob("*.cu"))


def get_file_hash(paths=None) -> str:
    """SHA-256 over every source, header and the flag set (reference: build.py:35-38 hashed one file)."""
    h = hashlib.sha256()
    files = list(paths) if paths is not None else sources() + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    for p in files:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and HASH_PATH.exists() and HASH_PATH.read_text().strip() == get_file_hash()


def build_cuda_kernels(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a and link the shared library.  Returns its path.

    Cross-compiles without a GPU.  Raises RuntimeError with the compiler output on failure.
    """
    force = force or os.environ.get("SD_B200_REBUILD", "0") == "1"
    if not force and is_current():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        if LIB_PATH.exists():
            logger.warning("nvcc not found; using the prebuilt %s (hash not verified)", LIB_PATH)
            return LIB_PATH
        raise RuntimeError("nvcc not found and no prebuilt libsd_verify_b200.so is present")
    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    srcs = sources()
    logs: Dict[str, str] = {}

    def compile_one(src: Path) -> Path:
        obj = obj_dir / (src.stem + ".o")
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src.name] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    link = [nvcc, *ARCH_FLAGS, "-shared", "-o", str(LIB_PATH), *map(str, objs)]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    HASH_PATH.write_text(get_file_hash())
    (LIB_DIR / "ptxas_info.log").write_text("\n".join(f"== {k}\n{v}" for k, v in sorted(logs.items())))
    if verbose:
        for k, v in sorted(logs.items()):
            print(f"== {k}\n{v}")
    logger.info("built %s", LIB_PATH)
    return LIB_PATH


# ---- the torch shim (csrc_torch/sd_torch_ops.cpp): host code only, g++ against the torch headers ------------------------
TORCH_SHIM_SRC = PKG_ROOT / "csrc_torch" / "sd_torch_ops.cpp"
TORCH_SHIM_PATH = LIB_DIR / "sd_b200_torch.so"
TORCH_SHIM_HASH = LIB_DIR / "sd_b200_torch.hash"


def _shim_hash() -> str:
    import torch

    h = hashlib.sha256()
    h.update(TORCH_SHIM_SRC.read_bytes())
    h.update((INCLUDE / "sd_verify.h").read_bytes())
    h.update(torch.__version__.encode())
    return h.hexdigest()


def build_torch_shim(force: bool = False) -> Path:
    """Compile the torch-side shim (argument marshalling + TORCH_LIBRARY registration; no kernels) and link it against
    libsd_verify_b200.so.  Ahead of time, in tree, hash-cached like the CUDA library: nothing is JIT-compiled at run
    time (a JIT cache under ~/.cache would not travel to the GPU box)."""
    force = force or os.environ.get("SD_B200_REBUILD", "0") == "1"
    if not force and TORCH_SHIM_PATH.exists() and TORCH_SHIM_HASH.exists() and TORCH_SHIM_HASH.read_text().strip() == _shim_hash():
        return TORCH_SHIM_PATH
    gxx = shutil.which("g++")
    if gxx is None:
        if TORCH_SHIM_PATH.exists():
            logger.warning("g++ not found; using the prebuilt %s (hash not verified)", TORCH_SHIM_PATH)
            return TORCH_SHIM_PATH
        raise RuntimeError("g++ not found and no prebuilt sd_b200_torch.so is present")
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce

    build_cuda_kernels()
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = [str(INCLUDE), *ce.include_paths(), sysconfig.get_paths()["include"], os.path.join(cuda_home, "include")]
    libdirs = [str(LIB_DIR), *ce.library_paths(), os.path.join(cuda_home, "lib64")]
    cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=sd_b200_torch", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for i in inc:
        cmd += ["-I", i]
    cmd += [str(TORCH_SHIM_SRC), "-o", str(TORCH_SHIM_PATH)]
    for d in libdirs:
        cmd += ["-L", d]
    cmd += ["-lsd_verify_b200", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10", "-lc10_cuda", "-lcudart", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for {TORCH_SHIM_SRC.name}:\n{r.stdout[-2000:]}\n{r.stderr[-6000:]}")
    TORCH_SHIM_HASH.write_text(_shim_hash())
    logger.info("built %s", TORCH_SHIM_PATH)
    return TORCH_SHIM_PATH


_SHIM = None


def load_torch_shim():
    """Import lib/sd_b200_torch.so (builds it when stale).  Importing it registers sd_b200::verify_prefix,
    sd_b200::verify_prefix_bonus and sd_b200::kv_append_rows_ with the torch dispatcher (CUDA key only)."""
    global _SHIM
    if _SHIM is None:
        import ctypes
        import importlib.util

        path = build_torch_shim()
        ctypes.CDLL(str(build_cuda_kernels()), mode=ctypes.RTLD_GLOBAL)  # (also found through the shim's rpath)
        spec = importlib.util.spec_from_file_location("sd_b200_torch", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _SHIM = mod
    return _SHIM


def load_kernels():
    """Build if needed and return the loaded C-ABI binding (kernels._capi.Lib).

    reference: src/kernels/build.py:136-187 returned a dict after walking a CUDA -> Triton -> Python
    fallback chain; here a failure to build or load is an error, never a fallback."""
    from . import _capi

    build_cuda_kernels()
    return _capi.lib()
