"""Build libgava_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m gava_clip_b200.build [--force] [--verbose]

The shared object lands in gava_clip_b200/lib/ (git-ignored; it travels with the repo snapshot).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libgava_sm100.so"
STAMP = LIBDIR / ".build_stamp"
SOURCES = ["api.cu", "gemm_tcgen05.cu", "rowwise.cu", "attention.cu", "attention_tc.cu", "precise.cu", "backward.cu",
           "attention_bwd.cu", "input_stage.cu", "seq_attention_mma.cu"]
HEADERS = ["common.cuh", "kernels.h", "mma_sync.cuh", "rowwise.cuh", "attention_tc.cuh", "../../include/gava_sm100.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--use_fast_math", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(name: str, defines) -> Path:
    """Experiment build (tools/gemm_bench.py): same sources with extra -D flags -> lib/variants/libgava_<name>.so."""
    vdir = LIBDIR / "variants"
    vdir.mkdir(parents=True, exist_ok=True)
    out = vdir / f"libgava_{name}.so"
    cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], *[f"-D{d}" for d in defines], "-shared",
           *[str(CSRC / s) for s in SOURCES], "-o", str(out), "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"variant build {name} failed")
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    objs = []
    for src in SOURCES:
        obj = LIBDIR / (src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        (LIBDIR / (src + ".ptxas.log")).write_text(r.stderr)
        objs.append(str(obj))
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
