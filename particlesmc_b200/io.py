"""Configuration files either side of the hot path: readers and writers for the reference's three on-disk formats
(src/IO/IO.jl, xyz.jl, exyz.jl, lammps.jl) and ``load_chains`` (IO.jl:208-330), which turns files into the
``Atoms`` / ``Molecules`` systems the device path consumes.

Host-side text handling only -- no arithmetic of the hot path lives here.  What is reproduced is the file grammar:

  XYZ     line 1 ``N``; line 2 ``step:t columns:[molecule,]species,position dt:1 cell:Lx,Ly[,Lz] rho:.. T:..``; then N
          rows ``[molecule] species x y [z]`` (xyz.jl:79-84); molecules append a bond table ``Nb`` / ``columns:bond`` /
          ``i j`` rows, 1-based (IO.jl:348-363);
  EXYZ    line 2 ``Lattice="Lx 0.0 0.0 0.0 Ly 0.0 0.0 0.0 Lz" Properties=[molecule:I:1]:species:S:1:pos:R:d Time=t``
          (exyz.jl:91-96); bond table header ``Properties=bond:I:2``;
  LAMMPS  dump: ``ITEM: TIMESTEP`` / t / ``ITEM: NUMBER OF ATOMS`` / N / ``ITEM: BOX BOUNDS pp pp pp`` / d bound
          rows (2-D adds ``-0.1 0.1``) / ``ITEM: ATOMS [molecule] type x y [z]`` (lammps.jl:88-105); no bond table.

Coordinates are written with 6 decimals by default (IO.jl:332-346, ``digits``).
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import models as M
from .systems import Atoms, EmptyList, LinkedList, Molecules, Particles, System, fold_back


class Format:
    extension = ""


class XYZ(Format):
    extension = ".xyz"


class EXYZ(Format):
    extension = ".exyz"


class LAMMPS(Format):
    extension = ".lammpstrj"


def _format_of(filename: str) -> Format:
    if filename.endswith(".xyz"):
        return XYZ()
    if filename.endswith(".exyz"):
        return EXYZ()
    if filename.endswith((".lmp", ".lammpstrj", ".lammps")):
        return LAMMPS()
    raise ValueError(f"Unsupported file format: {filename}")


# ---- headers ----------------------------------------------------------------------------------------------
def _columns_xyz(column_str: str, d: int) -> Dict[str, List[int]]:
    info, index = {}, 0
    for name in column_str.split(","):
        if name in ("molecule", "species", "btype"):
            info[name] = [1, index]
        elif name == "position":
            info["pos"] = [d, index]
        elif name == "bond":
            info["bond"] = [2, index]
        else:
            raise ValueError(f"{name} is not supported")
        index += 1  # the reference counts columns, not fields (xyz.jl:12-37); position is always last
    return info


def _columns_exyz(column_str: str) -> Dict[str, List[int]]:
    cols = column_str.split(":")
    info, i, index = {}, 0, 0
    while i < len(cols):
        if i + 2 < len(cols) and cols[i + 1] in ("S", "I", "R"):
            dim = int(cols[i + 2].split()[0])
            info[cols[i]] = [dim, index]
            index += dim
            i += 3
        else:
            i += 1
    return info


def _columns_lammps(line: str) -> Dict[str, List[int]]:
    cols = line.split()
    info, index = {}, 0
    for name in cols:
        if name in ("ITEM:", "ATOMS", "y", "z"):
            if name in ("y", "z"):
                index += 1
            continue
        if name == "molecule":
            info["molecule"] = [1, index]
        elif name == "type":
            info["species"] = [1, index]
        elif name == "x":
            info["pos"] = [3 if {"x", "y", "z"} <= set(cols) else 2, index]
        else:
            raise ValueError(f"{name} is not supported")
        index += 1
    return info


def _read_header(data: Sequence[str], fmt: Format):
    if isinstance(fmt, XYZ):
        N = int(data[0])
        metadata = data[1].split(" ")
        cell = next(t for t in metadata if t.startswith("cell:"))[len("cell:"):]
        box = np.array([float(v) for v in cell.split(",")])
        cols = next(t for t in metadata if t.startswith("columns:"))[len("columns:"):]
        return N, box, _columns_xyz(cols, len(box)), metadata, 2
    if isinstance(fmt, EXYZ):
        N = int(data[0])
        mat = re.search(r'Lattice="(.*?)"', data[1])
        if mat is None:
            raise ValueError("Invalid Lattice line format")
        vals = [float(v) for v in mat.group(1).split()]
        if len(vals) != 9:
            raise ValueError("Lattice matrix must have 9 elements")
        box = np.array([vals[0], vals[4], vals[8]])
        props = re.search(r"Properties=(\S*)", data[1]).group(1)
        return N, box, _columns_exyz(props), data[1].split(" "), 2
    n_idx = next(k for k, l in enumerate(data) if "ITEM: NUMBER OF ATOMS" in l)
    b_idx = next(k for k, l in enumerate(data) if "ITEM: BOX BOUNDS" in l)
    c_idx = next(k for k, l in enumerate(data) if "ITEM: ATOMS" in l)
    N = int(data[n_idx + 1])
    bounds = [[float(v) for v in data[b_idx + 1 + a].split()] for a in range(3)]
    box = np.array([hi - lo for lo, hi in bounds])
    return N, box, _columns_lammps(data[c_idx]), [], c_idx + 1


def _parse_label(tok: str):
    try:
        return int(tok)
    except ValueError:
        return tok


def load_configuration(filename_or_lines, fmt: Optional[Format] = None, m: int = 1) -> dict:
    """IO.jl:27-100: frame ``m`` (1-based) of a configuration file -> dict(N, d, box, species, position, metadata
    [, molecule, bond]); bond lists hold 1-based partner indices per site, as ``System`` expects."""
    if isinstance(filename_or_lines, str):
        fmt = fmt or _format_of(filename_or_lines)
        with open(filename_or_lines) as f:
            data = f.read().splitlines()
    else:
        data = list(filename_or_lines)
        if fmt is None:
            raise ValueError("a Format is needed when lines are passed")
    N, box, info, metadata, first = _read_header(data, fmt)
    stride = N + (9 if isinstance(fmt, LAMMPS) else 2)
    sel = first + stride * (m - 1)
    frame = data[sel:sel + N]
    if "pos" not in info:
        raise KeyError("pos array has not been found in metadata or is not defined. Define the pos in the args Dict")
    pos_d, pos_i = info["pos"]
    has_mol, has_sp = "molecule" in info, "species" in info
    if has_mol and m != 1:
        raise ValueError("For molecular systems the frame index has to be equal to 1")
    rows = [ln.split() for ln in frame]
    position = np.array([[float(v) for v in r[pos_i:pos_i + pos_d]] for r in rows], dtype=np.float64).reshape(N, pos_d)
    species = np.array([_parse_label(r[info["species"][1]]) for r in rows]) if has_sp else np.ones(N, dtype=np.int64)
    out = dict(N=N, d=pos_d, box=box[:pos_d].copy(), species=species, position=position, metadata=metadata)
    if has_mol:
        out["molecule"] = np.array([int(r[info["molecule"][1]]) for r in rows], dtype=np.int64)
        out["bond"] = _read_bonds(data[sel + N:], N, fmt)
    return out


def _read_bonds(lines: Sequence[str], N: int, fmt: Format) -> List[List[int]]:
    """IO.jl:158-199: ``Nb`` / header naming a ``bond`` column pair / Nb rows ``i j`` (1-based)."""
    if len(lines) == 0:
        raise ValueError("No bonds found in the file")
    nb = int(lines[0])
    if isinstance(fmt, EXYZ):
        info = _columns_exyz(re.search(r"Properties=(\S*)", lines[1]).group(1))
    else:
        cols = next(t for t in lines[1].split(" ") if t.startswith("columns:"))[len("columns:"):]
        info = _columns_xyz(cols, 3)
    if "bond" not in info:
        raise ValueError(f"Bond array is not written in the {type(fmt).__name__} file")
    if info["bond"][0] != 2:
        raise ValueError(f"Bond dimension must be 2. Found {info['bond'][0]}.")
    k = info["bond"][1]
    bond: List[List[int]] = [[] for _ in range(N)]
    for row in lines[2:2 + nb]:
        a, b = (int(v) for v in row.split()[k:k + 2])
        bond[a - 1].append(b)
        bond[b - 1].append(a)
    return bond


# ---- models from metadata / TOML tables ----------------------------------------------------------------------
def _get_model(data: dict, i: int, j: int):
    """IO.jl:129-156: one entry of the ``[model."i-j"]`` tables of params.toml (keys as the reference spells them)."""
    m = data[f"{i}-{j}" if i <= j else f"{j}-{i}"]
    opt = lambda **kw: {k: v for k, v in kw.items() if v is not None}
    if m["name"] == "GeneralKG":
        return M.GeneralKG(m["epsilon"], m["sigma"], m["k"], m["r0"],
                           **opt(rcut=m.get("rcut"), epsbond=m.get("epsilonbond"), sigmabond=m.get("sigmabond"),
                                 rcutbond=m.get("rcutbond")))
    if m["name"] == "SmoothLennardJones":
        return M.SmoothLennardJones(m["epsilon"], m["sigma"], **opt(rcut=m.get("rcut")))
    if m["name"] == "LennardJones":
        return M.LennardJones(m["epsilon"], m["sigma"], shift_potential=m.get("shift_potential", True),
                              **opt(rcut=m.get("rcut")))
    raise ValueError(f"Model {m['name']} is not implemented")


def _model_matrix(spec, n_species: int):
    if isinstance(spec, dict):
        return [[_get_model(spec, i, j) for j in range(1, n_species + 1)] for i in range(1, n_species + 1)]
    if isinstance(spec, str):
        name = spec.strip()
        if name.endswith("()"):
            name = name[:-2]
        if not re.fullmatch(r"[A-Za-z_]\w*", name) or not hasattr(M, name):
            raise ValueError(f"Model {spec} is not implemented")
        return getattr(M, name)()
    return spec  # already a model matrix


def volume_sphere(r: float, d: int) -> float:
    """src/utils.jl: d-dimensional ball."""
    return 4.0 / 3.0 * np.pi * r ** 3 if d == 3 else (np.pi * r ** 2 if d == 2 else 2.0 * r)


def load_chains(init_path: str, args: Optional[dict] = None, filename: str = "", verbose: bool = False,
                compute_energy: bool = True) -> List[Particles]:
    """IO.jl:208-330: one system per configuration file (times ``nsim`` replicas); ``args`` may override density,
    temperature, model, nsim, list_type exactly as the ``[system]`` table of params.toml does."""
    args = args or {}
    files: List[str] = []
    if os.path.isfile(init_path):
        files.append(init_path)
    elif os.path.isdir(init_path):
        for root, _, names in os.walk(init_path):
            files += [os.path.join(root, n) for n in sorted(names) if filename in n]
    if not files:
        raise FileNotFoundError(init_path)
    cfgs = [load_configuration(f) for f in files]
    N, d = cfgs[0]["N"], cfgs[0]["d"]
    assert all(c["N"] == N and c["d"] == d for c in cfgs)
    positions = [c["position"].copy() for c in cfgs]
    boxes = [c["box"].copy() for c in cfgs]
    densities = [N / float(np.prod(b)) for b in boxes]
    metas = [c["metadata"] for c in cfgs]

    def meta_value(meta, key):
        hit = [t for t in meta if key in t]
        return hit[0].split(":")[1] if hit else None

    temps = [meta_value(m, "T:") for m in metas]
    temperatures = [float(t) for t in temps] if all(t is not None for t in temps) else None
    mods = [meta_value(m, "model:") for m in metas]
    model_spec = mods[0] if all(t is not None for t in mods) else None
    if model_spec is not None:
        assert all(t == model_spec for t in mods)
    if args.get("density") is not None:
        for k in range(len(files)):
            lam = (densities[k] / args["density"]) ** (1.0 / d)
            positions[k] *= lam
            boxes[k] *= lam
            densities[k] = float(args["density"])
    if args.get("temperature") is not None:
        T = args["temperature"]
        temperatures = list(T) if isinstance(T, (list, tuple, np.ndarray)) else [float(T)] * len(files)
    elif temperatures is None:
        raise KeyError("temperature array has not been found in metadata or is not defined. Define the temperature in the args Dict")
    if args.get("model") is not None:
        model_spec = args["model"][0] if isinstance(args["model"], (list, tuple)) else args["model"]
    elif model_spec is None:
        raise KeyError("model array has not been found in metadata or is not defined. Define the model in the args Dict")
    positions = [fold_back(x, b) for x, b in zip(positions, boxes)]
    species = [c["species"] for c in cfgs]
    molecules = [c.get("molecule") for c in cfgs]
    bonds = [c.get("bond") for c in cfgs]
    nsim = args.get("nsim") or 1
    if nsim > 1:
        rep = lambda xs: [x if x is None else (x.copy() if hasattr(x, "copy") else x) for x in xs for _ in range(nsim)]
        positions, species, densities, temperatures = rep(positions), rep(species), rep(densities), rep(temperatures)
        molecules, bonds = rep(molecules), rep(bonds)
    n_species = len(np.unique(np.concatenate(species)))
    model_matrix = _model_matrix(model_spec, n_species)
    Z = float(np.mean(densities)) * volume_sphere(M.max_cutoff(model_matrix), d)
    list_type = LinkedList if Z / N < 0.1 else EmptyList  # IO.jl:308-310 (the device picks its own structure)
    if args.get("list_type") is not None:
        from . import systems as S
        list_type = getattr(S, str(args["list_type"]))
    chains = []
    for k in range(len(positions)):
        if molecules[k] is not None:
            chains.append(System(positions[k], species[k], molecules[k], densities[k], temperatures[k], model_matrix,
                                 bonds[k], list_type=list_type, compute_energy=compute_energy))
        else:
            chains.append(System(positions[k], species[k], densities[k], temperatures[k], model_matrix,
                                 list_type=list_type, compute_energy=compute_energy))
    if verbose:
        print(f"{len(chains)} chains created")
    return chains


# ---- writers -------------------------------------------------------------------------------------------------
def _num(v) -> str:
    return repr(float(v))


def write_header(io, system: Particles, t: int, fmt: Format, digits: int = 6):
    mol = isinstance(system, Molecules)
    if isinstance(fmt, XYZ):
        io.write(f"{len(system)}\n")
        box = ",".join(_num(v) for v in system.box)
        io.write(f"step:{t} columns:{'molecule,' if mol else ''}species,position dt:1 cell:{box} "
                 f"rho:{_num(system.density)} T:{_num(system.temperature)}\n")
    elif isinstance(fmt, EXYZ):
        io.write(f"{len(system)}\n")
        b = [_num(v) for v in system.box]
        lat = f"{b[0]} 0.0 0.0 0.0 {b[1]} 0.0 0.0 0.0 {b[2] if len(b) == 3 else '0.0'}"
        io.write(f"Lattice=\"{lat}\" Properties={'molecule:I:1' if mol else ''}:species:S:1:pos:R:{system.d} Time={t}\n")
    else:
        io.write(f"ITEM: TIMESTEP\n{t}\nITEM: NUMBER OF ATOMS\n{len(system)}\nITEM: BOX BOUNDS pp pp pp\n")
        for a in range(system.d):
            io.write(f"0.0 {_num(system.box[a])}\n")
        if system.d == 2:
            io.write("-0.1 0.1\n")
        io.write(f"ITEM: ATOMS {'molecule' if mol else ''} type x y{' z' if system.d == 3 else ''}\n")


def store_trajectory(io, system: Particles, t: int, fmt: Format, digits: int = 6):
    """IO.jl:365-381: header + one row per particle."""
    write_header(io, system, t, fmt, digits)
    mol = isinstance(system, Molecules)
    for k in range(len(system)):
        head = f"{system.molecule[k]} {system.species[k]}" if mol else f"{system.species[k]}"
        io.write(head + "".join(f" {v:.{digits}f}" for v in system.position[k]) + "\n")


def store_bonds(io, system: Molecules, fmt: Format):
    """IO.jl:348-363."""
    if isinstance(fmt, LAMMPS):
        raise ValueError("LAMMPS format does not support bonds format yet.")
    io.write(f"{sum(len(b) for b in system.bonds) // 2}\n")
    io.write("columns:bond\n" if isinstance(fmt, XYZ) else "Properties=bond:I:2\n")
    for i in range(1, system.N + 1):
        for j in system.bonds[i - 1]:
            if i < j:
                io.write(f"{i} {j}\n")


def store_lastframe(io, system: Particles, t: int, fmt: Format, digits: int = 6):
    """IO.jl:383-391: molecules carry their bond table in last-frame files."""
    store_trajectory(io, system, t, fmt, digits)
    if isinstance(system, Molecules):
        store_bonds(io, system, fmt)
