"""CPU tests of the drop-in boundary: the shared library builds for sm_100a, loads, and exports every symbol
include/pmc_b200.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from particlesmc_b200 import _lib as L
from particlesmc_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmc_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    path = B.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in pmc_b200.h but not exported"
    assert sorted(L.EXPORTS) == syms
    assert L.load().pmc_abi_version() == 1


def test_struct_layouts_match_header():
    assert ctypes.sizeof(L.Config) == 16 * 4
    assert ctypes.sizeof(L.MoveSpec) == 32
    assert ctypes.sizeof(L.Trial) == 48


def test_sass_is_sm100a():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", B.build_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_device_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from particlesmc_b200.device import DeviceContext
    with pytest.raises(L.PMCError, match="no CPU fallback"):
        DeviceContext(1, 10, 3, 1, 1)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "particlesmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle-free", ""), f"{f} mentions the oracle"


def _header_prototypes():
    """name -> number of parameters, from the declarations of include/pmc_b200.h."""
    text = open(os.path.join(ROOT, "include", "pmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(pmc_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def _ccalls(text):
    """(symbol, number of argument types) of every ccall((:pmc_x, LIB), Ret, (T1, T2, ...), ...) in a Julia source."""
    found = []
    for m in re.finditer(r"ccall\(\(:(pmc_[a-z0-9_]+),\s*LIB\),\s*\w+,\s*\(", text):
        i, depth, start = m.end(), 1, m.end()
        while depth:  # the type tuple may nest braces / parentheses (Ptr{Cvoid}, Ref{Ptr{Cvoid}})
            depth += text[i] in "({"
            depth -= text[i] in ")}"
            i += 1
        types = text[start:i - 1]
        parts, cur, d = [], "", 0
        for ch in types:
            if ch in "({":
                d += 1
            if ch in ")}":
                d -= 1
            if ch == "," and d == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        parts.append(cur)
        found.append((m.group(1), len([p for p in parts if p.strip()])))
    return found


def test_julia_shim_and_integration_guide_bind_the_declared_prototypes():
    """The reference-side binding (julia/*.jl, the stubs of INTEGRATION.md) must call what the header declares: every
    symbol exists and every ccall passes as many arguments as the C prototype takes.  Julia is not in this image, so
    this is the check that keeps the shim from drifting when the ABI changes."""
    protos = _header_prototypes()
    assert len(protos) >= 25
    n_calls = 0
    for rel in ("julia/ParticlesMCB200.jl", "julia/replay_check.jl", "INTEGRATION.md"):
        text = open(os.path.join(ROOT, rel)).read()
        for name, nargs in _ccalls(text):
            assert name in protos, f"{rel}: {name} is not declared in pmc_b200.h"
            assert nargs == protos[name], f"{rel}: {name} bound with {nargs} arguments, the header declares {protos[name]}"
            n_calls += 1
    assert n_calls >= 25
