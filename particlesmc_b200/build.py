"""Builds the C-ABI shared library ``lib/libpmc_b200.so`` from ``csrc/*.cu`` with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting ``.so`` is
git-ignored but travels to the GPU box inside the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libpmc_b200.so")
SOURCES = ["api.cu", "chains.cu", "chains_fast.cu", "chains_spec.cu", "box.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
]


def _newest_source_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "pmc_b200.h")]
    return max(os.path.getmtime(p) for p in paths)


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libpmc_b200.so")
    return nvcc


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile if the library is missing or older than any source. Returns the library path."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    import tempfile
    obj_dir = os.path.join(tempfile.gettempdir(), "pmc_b200_build")  # objects stay out of the repo snapshot
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = find_nvcc()
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]

    def compile_one(src: str):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *inc, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, res

    # the translation units are independent: compile them concurrently, then link
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for src, _, res in results:
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    link = subprocess.run([nvcc, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB_PATH, *[o for _, o, _ in results]],
                          capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("link failed:\n" + link.stdout + link.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
