"""Pair-potential models and named mixtures (host-side mirror of the reference's ``src/models.jl``).

Same names, constructor arguments and derived quantities as the reference so that a model
matrix built here flattens to exactly the numbers the Julia structs hold:

* ``SoftSpheres``          <- models.jl:52-74   (``inverse_power`` :28)
* ``LennardJones``         <- models.jl:99-123  (``lennard_jones`` :30-34)
* ``SmoothLennardJones``   <- models.jl:137-166
* ``GeneralKG``            <- models.jl:183-226 (``fene`` :36)
* ``BHHP() KobAndersen() JBB() Trimer()`` <- models.jl:76-84, :125-133, :168-179, :231-243

The device library never sees these objects: ``flatten_model_matrix`` turns an ``ns x ns``
matrix into the ``[ns][ns][PMC_NPAR]`` float64 block declared in ``include/pmc_b200.h``.
All derived constants are evaluated in IEEE double with the reference's operation order.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

PMC_NPAR = 12
MODEL_LJ, MODEL_SOFT, MODEL_SMOOTHLJ, MODEL_KG = 1, 2, 3, 4

# parameter slots (include/pmc_b200.h)
P_RCUT, P_RCUT2, P_EPS, P_SIG2, P_SHIFT = 0, 1, 2, 3, 4


def lennard_jones(r2: float, eps4: float, sig2: float) -> float:
    """models.jl:30-34."""
    x = sig2 * (1.0 / r2)
    x3 = x * x * x
    return eps4 * (x3 * x3 - x3)


def inverse_power(r2: float, eps: float, sig2: float, ndiv2) -> float:
    """models.jl:28."""
    return eps * (sig2 / r2) ** ndiv2


class Model:
    """abstract type Model (models.jl:8)."""

    kind: int = 0
    rcut: float
    rcut2: float

    def flat(self) -> np.ndarray:  # pragma: no cover - overridden
        raise NotImplementedError


@dataclass
class SoftSpheres(Model):
    """SoftSpheres(eps, sigma, n; rcut=2.5*sigma) -- models.jl:52-70."""

    eps: float
    sigma: float
    n: int
    rcut: float | None = None
    name: str = "SoftSpheres"
    kind = MODEL_SOFT

    def __post_init__(self):
        if self.rcut is None:
            self.rcut = 2.5 * self.sigma
        self.sig2 = self.sigma * self.sigma
        self.rcut2 = self.rcut * self.rcut
        self.ndiv2 = self.n / 2 if self.n % 2 else self.n // 2
        self.shift = inverse_power(self.rcut2, self.eps, self.sig2, self.ndiv2)

    def potential(self, r2: float) -> float:
        return inverse_power(r2, self.eps, self.sig2, self.ndiv2) - self.shift

    def flat(self) -> np.ndarray:
        p = np.zeros(PMC_NPAR)
        p[:6] = [self.rcut, self.rcut2, self.eps, self.sig2, self.shift, float(self.ndiv2)]
        return p


@dataclass
class LennardJones(Model):
    """LennardJones(eps, sigma; rcut=2.5*sigma, shift_potential=true) -- models.jl:99-119."""

    eps: float
    sigma: float
    rcut: float | None = None
    name: str = "LennardJones"
    shift_potential: bool = True
    kind = MODEL_LJ

    def __post_init__(self):
        if self.rcut is None:
            self.rcut = 2.5 * self.sigma
        self.eps4 = 4 * self.eps
        self.sig2 = self.sigma * self.sigma
        self.rcut2 = self.rcut * self.rcut
        self.shift = lennard_jones(self.rcut2, self.eps4, self.sig2) if self.shift_potential else 0.0

    def potential(self, r2: float) -> float:
        return lennard_jones(r2, self.eps4, self.sig2) - self.shift

    def flat(self) -> np.ndarray:
        p = np.zeros(PMC_NPAR)
        p[:5] = [self.rcut, self.rcut2, self.eps4, self.sig2, self.shift]
        return p


@dataclass
class SmoothLennardJones(Model):
    """SmoothLennardJones(eps, sigma; rcut=2.5*sigma) -- models.jl:137-158."""

    eps: float
    sigma: float
    rcut: float | None = None
    name: str = "SmoothLennardJones"
    kind = MODEL_SMOOTHLJ

    def __post_init__(self):
        C0, C2, C4 = 0.04049023795, -0.00970155098, 0.00062012616
        if self.rcut is None:
            self.rcut = 2.5 * self.sigma
        self.eps4 = 4 * self.eps
        self.sig2 = self.sigma * self.sigma
        self.C0 = C0
        self.C2_sig2 = C2 / self.sig2
        self.C4_sig4 = C4 / (self.sig2 * self.sig2)
        self.rcut2 = self.rcut * self.rcut

    def potential(self, r2: float) -> float:
        lj = lennard_jones(r2, self.eps4, self.sig2)
        return lj + self.eps4 * (self.C0 + r2 * (r2 * self.C4_sig4 + self.C2_sig2))

    def flat(self) -> np.ndarray:
        p = np.zeros(PMC_NPAR)
        p[:8] = [self.rcut, self.rcut2, self.eps4, self.sig2, 0.0, self.C0, self.C2_sig2, self.C4_sig4]
        return p


@dataclass
class GeneralKG(Model):
    """GeneralKG(eps, sigma, k, r0; rcut=2^(1/6)*sigma, epsbond=eps, sigmabond=sigma, rcutbond=rcut)
    -- models.jl:183-217."""

    eps: float
    sigma: float
    k: float
    r0: float
    rcut: float | None = None
    epsbond: float | None = None
    sigmabond: float | None = None
    rcutbond: float | None = None
    name: str = "GeneralKG"
    kind = MODEL_KG

    def __post_init__(self):
        if self.rcut is None:
            self.rcut = 2 ** (1 / 6) * self.sigma
        if self.epsbond is None:
            self.epsbond = self.eps
        if self.sigmabond is None:
            self.sigmabond = self.sigma
        if self.rcutbond is None:
            self.rcutbond = self.rcut
        self.r02 = self.r0 * self.r0
        self.rcut2 = self.rcut * self.rcut
        self.rcut2bond = self.rcutbond * self.rcutbond
        self.kr02 = -self.k * self.r02 / 2
        self.eps4 = 4 * self.eps
        self.eps4bond = 4 * self.epsbond
        self.sig2 = self.sigma * self.sigma
        self.sig2bond = self.sigmabond * self.sigmabond
        self.shift = lennard_jones(self.rcut2, self.eps4, self.sig2)
        self.shiftbond = lennard_jones(self.rcut2bond, self.eps4bond, self.sig2bond)

    def potential(self, r2: float) -> float:
        return lennard_jones(r2, self.eps4, self.sig2) - self.shift

    def bond_potential(self, r2: float) -> float:
        u_fene = self.kr02 * math.log(1 - r2 * (1.0 / self.r02)) if r2 <= self.r02 else math.inf
        u_lj = 0.0
        if r2 <= self.rcut2bond:
            u_lj += lennard_jones(r2, self.eps4bond, self.sig2bond) - self.shiftbond
        return u_fene + u_lj

    def flat(self) -> np.ndarray:
        p = np.zeros(PMC_NPAR)
        p[:11] = [self.rcut, self.rcut2, self.eps4, self.sig2, self.shift, self.eps4bond, self.sig2bond,
                  self.rcut2bond, self.shiftbond, self.kr02, self.r02]
        return p


def cutoff(model: Model) -> float:
    return model.rcut


def cutoff2(model: Model) -> float:
    return model.rcut2


def _matrix(ctor, eps, sig, *extra) -> List[List[Model]]:
    n = len(eps)
    return [[ctor(eps[i][j], sig[i][j], *[e[i][j] for e in extra]) for j in range(n)] for i in range(n)]


def BHHP() -> List[List[Model]]:
    """models.jl:76-84: binary soft spheres, n = 12."""
    eps = [[1.0, 1.0], [1.0, 1.0]]
    sig = [[1.0, 1.2], [1.2, 1.4]]
    return [[SoftSpheres(eps[i][j], sig[i][j], 12) for j in range(2)] for i in range(2)]


def KobAndersen() -> List[List[Model]]:
    """models.jl:125-133: the 80:20 Kob-Andersen Lennard-Jones mixture."""
    return _matrix(LennardJones, [[1.0, 1.5], [1.5, 0.5]], [[1.0, 0.8], [0.8, 0.88]])


def JBB() -> List[List[Model]]:
    """models.jl:168-179: ternary smooth-LJ mixture."""
    eps = [[1.0, 1.5, 0.75], [1.5, 0.5, 1.5], [0.75, 1.5, 0.75]]
    sig = [[1.0, 0.8, 0.9], [0.8, 0.88, 0.8], [0.9, 0.8, 0.94]]
    return _matrix(SmoothLennardJones, eps, sig)


def Trimer() -> List[List[Model]]:
    """models.jl:231-243: three-site molecule, WCA non-bonded + FENE bonds."""
    eps = [[1.0] * 3] * 3
    sig = [[0.9, 0.95, 1.0], [0.95, 1.0, 1.05], [1.0, 1.05, 1.1]]
    k = [[0.0, 33.241, 30.0], [33.241, 0.0, 27.210884], [30.0, 27.210884, 0.0]]
    r0 = [[0.0, 1.425, 1.5], [1.425, 0.0, 1.575], [1.5, 1.575, 0.0]]
    return _matrix(GeneralKG, eps, sig, k, r0)


NAMED_MODELS = {"BHHP": BHHP, "KobAndersen": KobAndersen, "JBB": JBB, "Trimer": Trimer}


def get_model(table: dict) -> Model:
    """TOML ``[model."i-j"]`` table -> model (IO.jl:129-156)."""
    name = table["name"]
    if name == "GeneralKG":
        # TOML keys of the reference schema (IO.jl:136-141): epsilonbond / sigmabond / rcutbond
        kw = {k: table[k] for k in ("rcut", "sigmabond", "rcutbond") if k in table}
        if "epsilonbond" in table:
            kw["epsbond"] = table["epsilonbond"]
        return GeneralKG(table["epsilon"], table["sigma"], table["k"], table["r0"], **kw)
    if name == "SmoothLennardJones":
        return SmoothLennardJones(table["epsilon"], table["sigma"], rcut=table.get("rcut"))
    if name == "LennardJones":
        return LennardJones(table["epsilon"], table["sigma"], rcut=table.get("rcut"),
                            shift_potential=table.get("shift_potential", True))
    raise ValueError(f"Model {name} is not implemented")


def model_kind(model_matrix: Sequence[Sequence[Model]]) -> int:
    kinds = {m.kind for row in model_matrix for m in row}
    if len(kinds) != 1:
        raise ValueError("all entries of a model matrix must be of the same model type")
    return kinds.pop()


def flatten_model_matrix(model_matrix: Sequence[Sequence[Model]]) -> np.ndarray:
    """``ns x ns`` model matrix -> C-contiguous float64 ``[ns][ns][PMC_NPAR]`` block."""
    ns = len(model_matrix)
    out = np.zeros((ns, ns, PMC_NPAR), dtype=np.float64)
    for i in range(ns):
        if len(model_matrix[i]) != ns:
            raise ValueError("model matrix must be square")
        for j in range(ns):
            out[i, j] = model_matrix[i][j].flat()
    return out


def max_cutoff(model_matrix: Sequence[Sequence[Model]]) -> float:
    """maximum([model.rcut for model in model_matrix]) (atoms.jl:46)."""
    return max(m.rcut for row in model_matrix for m in row)
