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
# PMC_B200_LIB names an alternative build of the library (A/B measurements of build-time kernel switches, built with
# `python -m particlesmc_b200.build --out NAME -DSWITCH=VALUE ...`); the product default is libpmc_b200.so
LIB_PATH = os.path.join(LIB_DIR, os.environ.get("PMC_B200_LIB", "libpmc_b200.so"))
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


def build_library(force: bool = False, verbose: bool = False, out: str | None = None, defines: tuple = ()) -> str:
    """Compile if the library is missing or older than any source. Returns the library path."""
    lib_path = os.path.join(LIB_DIR, out) if out else LIB_PATH
    if not force and os.path.exists(lib_path) and os.path.getmtime(lib_path) >= _newest_source_mtime():
        return lib_path
    os.makedirs(LIB_DIR, exist_ok=True)
    import tempfile
    # objects stay out of the repo snapshot
    obj_dir = os.path.join(tempfile.gettempdir(), "pmc_b200_build_" + os.path.basename(lib_path))
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = find_nvcc()
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]

    def compile_one(src: str):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, *inc, "-c", os.path.join(CSRC, src), "-o", obj]
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
    link = subprocess.run([nvcc, "-shared", "-ccbin", "/usr/bin/g++", "-o", lib_path, *[o for _, o, _ in results]],
                          capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("link failed:\n" + link.stdout + link.stderr)
    return lib_path


if __name__ == "__main__":
    import sys

    out_name = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build_library(force="--force" in sys.argv or out_name is not None, verbose="-v" in sys.argv, out=out_name,
                        defines=tuple(a for a in sys.argv[1:] if a.startswith("-D"))))
