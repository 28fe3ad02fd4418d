"""Build ``libsdt_b200.so`` (the CUDA hot path + C ABI) in-tree with nvcc for sm_100a.

The library is the product path; nothing here falls back to another backend.  ``nvcc`` cross-compiles
without a GPU, so the same command serves the CPU-only build check and the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = BUILD_DIR / "libsdt_b200.so"
INCLUDE_DIR = PKG_DIR.parent / "include"

SOURCES = ["api.cu", "elementwise.cu", "groupnorm.cu", "layernorm.cu", "comm.cu", "simt_gemm.cu", "lora_gemm.cu", "lora_gemm2.cu", "lora_wgrad.cu", "lora_api.cu", "dropout.cu"]

COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]
NVCC_FLAGS = COMPILE_FLAGS + ["-shared"]      # kept for tools that compile a single file against the same flags


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libsdt_b200.so cannot be built (there is no non-CUDA path)")
    return cand


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob("*.cuh")) + list(INCLUDE_DIR.glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def _compile_one(nvcc: str, src: Path, obj: Path, verbose: bool) -> str:
    cmd = [nvcc, *COMPILE_FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n" + proc.stdout + proc.stderr)
    return proc.stderr


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if it is missing or older than its sources; return its path.  Every ``.cu`` becomes its own
    object (compiled in parallel, rebuilt only when it or a header is newer), then one link."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = BUILD_DIR / "obj"
    obj_dir.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    headers = list(CSRC.glob("*.cuh")) + list(INCLUDE_DIR.glob("*.h"))
    t_hdr = max(h.stat().st_mtime for h in headers)
    jobs = []
    for s in SOURCES:
        src, obj = CSRC / s, obj_dir / (s[:-3] + ".o")
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, t_hdr):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1) or 1) as pool:
        logs = list(pool.map(lambda j: _compile_one(nvcc, j[0], j[1], verbose), jobs))
    if verbose:
        print("\n".join(logs))
    tmp = BUILD_DIR / "libsdt_b200.so.tmp"
    cmd = [nvcc, *LINK_FLAGS, *(str(obj_dir / (s[:-3] + ".o")) for s in SOURCES), "-o", str(tmp)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
