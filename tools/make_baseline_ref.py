#!/usr/bin/env python3
"""Copy the UNMODIFIED reference sources into baseline/_ref/ (git-ignored, NOT gpurun-ignored).

    python tools/make_baseline_ref.py [--reference /root/reference]

The reference (GogoRit/llm-inference-lab) is plain Python: nothing is built.  baseline/_ref/src is what
`bench.py --impl reference`, the `cpu_baseline` leg and tests/test_reference_callers_gpu.py import -- the reference
itself as the reference arm, and the reference's own callers (policies, HFWrapper, SpeculativePipeline) driving
this package's kernels.  /root/reference does not exist on the GPU box; this copy travels with the repo snapshot.
A manifest with the SHA-256 of every copied file is written next to it so a reader can check nothing was edited.
Never commit baseline/_ref (it is the reference's code, not this repository's)."""
import argparse
import hashlib
import json
import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    src = Path(a.reference)
    dst = ROOT / "baseline" / "_ref"
    if not (src / "src" / "kernels" / "reference.py").exists():
        print(f"{src}: not a reference checkout; keeping the existing {dst}" if dst.exists() else f"{src}: not found")
        return 0 if dst.exists() else 1
    if dst.exists():
        shutil.rmtree(dst)
    dst.mkdir(parents=True)
    manifest = {}
    for sub in ("src", "scripts", "configs"):
        if not (src / sub).exists():
            continue
        for p in sorted((src / sub).rglob("*")):
            if p.is_dir() or "__pycache__" in p.parts or p.suffix in (".pyc", ".so"):
                continue
            rel = p.relative_to(src)
            out = dst / rel
            out.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(p, out)
            manifest[str(rel)] = hashlib.sha256(p.read_bytes()).hexdigest()
    (dst / "MANIFEST.json").write_text(json.dumps(manifest, indent=0, sort_keys=True))
    print(f"copied {len(manifest)} files to {dst}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
