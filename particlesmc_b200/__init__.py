"""particlesmc_b200 -- B200-native Metropolis hot path of ParticlesMC behind a C ABI.

Host-side mirror of the reference's public names for that path (src/ParticlesMC.jl:115-127).  The CUDA
library (``lib/libpmc_b200.so``, built from ``csrc/`` by ``particlesmc_b200.build``) does all the work;
nothing here computes energies or moves on the CPU.
"""
from .models import (BHHP, JBB, GeneralKG, KobAndersen, LennardJones, Model, SmoothLennardJones, SoftSpheres, Trimer,
                     cutoff, cutoff2, flatten_model_matrix, get_model, model_kind)
from .moves import (Action, DiscreteSwap, Displacement, DoubleUniform, MoleculeFlip, Move, Policy, SimpleGaussian,
                    delta_log_target_density, log_proposal_density)
from .systems import (Atoms, CellList, EmptyList, LinkedList, Molecules, NeighbourList, Particles, System, VerletList,
                      bonds_from_pairs, compute_energy_particle, energy, fold_back, make_context)
from .simulation import (Metropolis, PrintTimeSteps, Simulation, StoreAcceptance, StoreCallbacks, StoreLastFrames,
                         StoreTrajectories, build_schedule, run)
from .device import DeviceContext, measure_fma_peak
from .observables import EnergyHistogram, chain_correlation, radial_distribution
from .io import EXYZ, LAMMPS, XYZ, load_chains, load_configuration, store_lastframe, store_trajectory
from ._lib import PMCError

__all__ = [n for n in dir() if not n.startswith("_")]
