"""The reference's integration tests (test/runtests.jl:40-129, 151-191) through the host mirror: build systems with
System(...), run a Simulation with the Metropolis entry + output algorithms, read chains/<k>/energy.dat back.
Where the reference compares EmptyList against LinkedList, the device compares its two candidate-visiting
strategies (every candidate in fp64 vs the fixed-point prefilter), which must give the same energy series."""
import os

import numpy as np
import pytest

import particlesmc_b200 as P
from particlesmc_b200.device import DeviceContext

pytestmark = pytest.mark.gpu


def read_energy(path):
    return np.loadtxt(os.path.join(path, "chains", "1", "energy.dat"))


def make_algorithms(pool, seed, N, steps, sampletimes):
    return (
        dict(algorithm=P.Metropolis, pool=pool, seed=seed, parallel=False, sweepstep=N),
        dict(algorithm=P.StoreCallbacks, callbacks=(P.energy,), scheduler=sampletimes),
        dict(algorithm=P.StoreAcceptance, dependencies=(P.Metropolis,), scheduler=sampletimes),
        dict(algorithm=P.StoreTrajectories, scheduler=[0, steps]),
        dict(algorithm=P.StoreLastFrames, scheduler=[steps]),
        dict(algorithm=P.PrintTimeSteps, scheduler=P.build_schedule(steps, 0, steps // 10)),
    )


def test_atoms_simulation_with_and_without_swaps(config0, tmp_path):
    system = P.System(config0["position"], config0["species"], config0["density"], config0["temperature"], P.JBB(),
                      list_type=P.LinkedList)
    assert abs(P.energy(system) - (-2.676832)) < 1e-6
    NA, NB, NC = (int(np.count_nonzero(system.species == s)) for s in (1, 2, 3))
    steps, seed = 20, 10
    sampletimes = P.build_schedule(steps, 0, [0, 1, 2, 4, 8])
    for pswap in (0.0, 0.8):
        pool = [P.Move(P.Displacement(0, np.zeros(2), 0.0), P.SimpleGaussian(), {"sigma": 0.05}, 1 - pswap)]
        if pswap:
            pool += [P.Move(P.DiscreteSwap(0, 0, (1, 3), (NA, NC), 0.0), P.DoubleUniform(), [], pswap / 2),
                     P.Move(P.DiscreteSwap(0, 0, (2, 3), (NB, NC), 0.0), P.DoubleUniform(), [], pswap / 2)]
        series = []
        for tag, prefilter in (("prefilter", 0), ("all_candidates", -1)):
            import copy
            chains = [copy.deepcopy(system)]
            path = str(tmp_path / f"swap{pswap}" / tag)
            sim = P.Simulation(chains, make_algorithms(pool, seed, system.N, steps, sampletimes), steps, path=path)
            # the prefilter switch lives in pmc_config: rebuild the context of this simulation accordingly
            if prefilter:
                sim.ctx.close()
                from particlesmc_b200.systems import make_context
                from particlesmc_b200.moves import pool_to_specs
                s0 = chains[0]
                ctx = DeviceContext(1, s0.N, s0.d, 3, P.model_kind(s0.model_matrix), prefilter=-1)
                ctx.set_model(P.flatten_model_matrix(s0.model_matrix))
                ctx.upload(s0.position, s0.species, s0.box, s0.temperature)
                ctx.init_energy()
                ctx.set_moves(pool_to_specs(pool))
                ctx.seed(seed)
                sim.ctx = ctx
            P.run(sim)
            e = read_energy(path)
            assert e[:, 0].tolist() == [float(t) for t in sampletimes]
            assert e[0, 1] == pytest.approx(-2.676832, abs=1e-6)
            # host copies are current at the end: energy[1], positions, counters
            assert P.energy(chains[0]) == pytest.approx(e[-1, 1], abs=1e-12)
            assert abs(sim.ctx.total_energy()[0] / system.N - e[-1, 1]) < 1e-10
            assert sum(mv.total_calls for mv in pool) == steps * system.N
            assert os.path.exists(os.path.join(path, "moves", "1", "acceptance.dat"))
            assert os.path.exists(os.path.join(path, "chains", "1", "lastframe.xyz"))
            traj = open(os.path.join(path, "chains", "1", "trajectory.xyz")).read().split("\n")
            assert traj[0] == "1290" and traj.count("1290") == 2
            series.append(e[:, 1])
            sim.close()
        assert np.allclose(series[0], series[1], atol=1e-6, rtol=0)  # the reference's list-equivalence criterion
        assert np.max(np.abs(series[0] - series[1])) < 1e-10


def test_molecules_simulation(molecule, tmp_path):
    system = P.System(molecule["position"], molecule["species"], molecule["molecule"], molecule["density"],
                      molecule["temperature"], P.Trimer(), molecule["bonds"], list_type=P.LinkedList)
    assert abs(P.energy(system) - 25.65865662277199) < 1e-6
    steps = 10
    pool = [P.Move(P.Displacement(0, np.zeros(3), 0.0), P.SimpleGaussian(), {"sigma": 0.05}, 1.0)]
    sampletimes = P.build_schedule(steps, 0, [0, 1, 2, 4, 8])
    path = str(tmp_path / "mol")
    sim = P.Simulation([system], make_algorithms(pool, 10, system.N, steps, sampletimes), steps, path=path)
    P.run(sim)
    e = read_energy(path)
    assert e[0, 1] == pytest.approx(25.65865662277199, abs=1e-6)
    assert abs(e[-1, 1] - e[0, 1]) < 1.0 and e[-1, 1] != e[0, 1]  # an equilibrated T=2 configuration: fluctuates
    assert abs(sim.ctx.total_energy()[0] / system.N - e[-1, 1]) < 1e-10
    assert pool[0].total_calls == steps * system.N and 0 < pool[0].accepted_calls < pool[0].total_calls
    sim.close()


def test_initial_overlap_raises_in_simulation(config0):
    pos = config0["position"].copy()
    pos[3] = pos[11]
    s = P.System(pos, config0["species"], config0["density"], config0["temperature"], P.JBB(), compute_energy=False)
    pool = [P.Move(P.Displacement(0, np.zeros(2), 0.0), P.SimpleGaussian(), {"sigma": 0.05}, 1.0)]
    with pytest.raises(ValueError, match="infinite or NaN energy"):
        P.Simulation([s], [dict(algorithm=P.Metropolis, pool=pool, seed=1, parallel=False, sweepstep=s.N)], 1)
