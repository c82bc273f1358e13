"""Action / Policy objects of the move pool: host-side mirror of ``src/moves.jl``.

In the reference every trial goes through ``sample_action! -> perform_action! -> revert_action!`` on the
host (src/moves.jl:57-90, :159-214).  Here these objects only DESCRIBE the pool; the trials themselves
run inside the sweep kernels.  ``pool_to_specs`` flattens a pool into the ``pmc_move`` records of the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence, Tuple

import numpy as np


class Action:
    pass


class Policy:
    pass


@dataclass
class Displacement(Action):
    """mutable struct Displacement (src/moves.jl:34-38): particle i moved by delta, energy change de."""
    i: int = 0
    delta: Any = None
    de: float = 0.0


@dataclass
class DiscreteSwap(Action):
    """mutable struct DiscreteSwap (src/moves.jl:137-143)."""
    i: int = 0
    j: int = 0
    species: Tuple[int, int] = (1, 2)
    particles_per_species: Tuple[int, int] = (0, 0)
    de: float = 0.0

    @classmethod
    def from_system(cls, species: Sequence[int], system) -> "DiscreteSwap":
        """DiscreteSwap(species::Vector{Int}, system::Atoms) (src/moves.jl:145-149)."""
        n1 = int(np.count_nonzero(system.species == species[0]))
        n2 = int(np.count_nonzero(system.species == species[1]))
        return cls(1, 1, (int(species[0]), int(species[1])), (n1, n2), 0.0)


@dataclass
class MoleculeFlip(Action):
    """mutable struct MoleculeFlip (src/moves.jl:291-295): species exchange between two unlike sites of a molecule."""
    i: int = 0
    j: int = 0
    de: float = 0.0


class SimpleGaussian(Policy):
    """struct SimpleGaussian <: Policy (src/moves.jl:105); parameters: sigma."""


class DoubleUniform(Policy):
    """struct DoubleUniform <: Policy (src/moves.jl:226)."""


@dataclass
class Move:
    """Arianna ``Move(action, policy, parameters, probability)`` as built at src/ParticlesMC.jl:192-245."""
    action: Action
    policy: Policy
    parameters: Any
    probability: float
    total_calls: int = 0
    accepted_calls: int = 0
    _per_chain: Any = field(default=None, repr=False)

    @property
    def sigma(self) -> float:
        p = self.parameters
        if isinstance(p, dict):
            return float(p.get("sigma", p.get("σ")))
        for name in ("sigma", "σ"):
            if hasattr(p, name):
                return float(getattr(p, name))
        return float(p)


def log_proposal_density(action: Action, policy: Policy, parameters, system) -> float:
    """src/moves.jl:110-112 (SimpleGaussian) and :231-233 (DoubleUniform)."""
    if isinstance(action, Displacement):
        sigma = parameters["sigma"] if isinstance(parameters, dict) else float(parameters)
        delta = np.asarray(action.delta, dtype=np.float64)
        return float(-np.dot(delta, delta) / (2 * sigma ** 2) - system.d * np.log(2 * np.pi * sigma ** 2) / 2)
    if isinstance(action, DiscreteSwap):
        return float(-np.log(action.particles_per_species[0] * action.particles_per_species[1]))
    if isinstance(action, MoleculeFlip):
        return float(-np.log(2))  # src/moves.jl:336-338
    raise TypeError(type(action))


def delta_log_target_density(e1: float, e2: float, system) -> float:
    """-(e2 - e1) / T (src/utils.jl:8-10)."""
    return -(e2 - e1) / system.temperature


def pool_to_specs(pool: Sequence[Move]):
    specs = []
    for mv in pool:
        if isinstance(mv.action, Displacement):
            if not isinstance(mv.policy, SimpleGaussian):
                raise NotImplementedError("Displacement is implemented with the SimpleGaussian policy")
            specs.append({"kind": "displacement", "prob": mv.probability, "sigma": mv.sigma})
        elif isinstance(mv.action, DiscreteSwap):
            if not isinstance(mv.policy, DoubleUniform):
                raise NotImplementedError("DiscreteSwap is implemented with the DoubleUniform policy "
                                          "(EnergyBias is policy-guided MC, outside the hot path)")
            specs.append({"kind": "swap", "prob": mv.probability, "species": tuple(mv.action.species)})
        elif isinstance(mv.action, MoleculeFlip):
            if not isinstance(mv.policy, DoubleUniform):
                raise NotImplementedError("MoleculeFlip is implemented with the DoubleUniform policy")
            specs.append({"kind": "flip", "prob": mv.probability})
        else:
            raise NotImplementedError(f"action {type(mv.action).__name__} is not on the device path")
    return specs
