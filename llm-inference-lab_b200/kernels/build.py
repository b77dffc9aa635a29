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


def load_kernels():
    """Build if needed and return the loaded C-ABI binding (kernels._capi.Lib).

    reference: src/kernels/build.py:136-187 returned a dict after walking a CUDA -> Triton -> Python
    fallback chain; here a failure to build or load is an error, never a fallback."""
    from . import _capi

    build_cuda_kernels()
    return _capi.lib()
