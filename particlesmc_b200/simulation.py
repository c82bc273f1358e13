"""Driver level: the ``Metropolis`` algorithm entry and a minimal ``Simulation`` / ``run`` around it.

In the reference this level belongs to Arianna.jl (external): ``Simulation(chains, algorithm_list, steps;
path)`` and ``run!`` (src/ParticlesMC.jl:294-297), the algorithm tuple ``(algorithm=Metropolis, pool, seed,
parallel, sweepstep)`` (src/ParticlesMC.jl:246), ``build_schedule`` (:255-261) and the output algorithms
StoreCallbacks / StoreAcceptance / StoreTrajectories / StoreLastFrames / PrintTimeSteps (:249-291).
This module provides the same call shapes so that the parity tests read like test/runtests.jl; one
Arianna step advances every chain by ``sweepstep`` trials in ONE device launch, and host copies of
position / species / energy[1] / move counters are refreshed only when an output schedule fires.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .moves import Move, pool_to_specs
from .systems import Particles, make_context


class Metropolis:
    """Marker for the algorithm entry replaced by the device sweep."""


class StoreCallbacks:
    pass


class StoreAcceptance:
    pass


class StoreTrajectories:
    pass


class StoreLastFrames:
    pass


class PrintTimeSteps:
    pass


def build_schedule(steps: int, burn: int, block) -> List[int]:
    """``build_schedule(steps, burn, interval::Int)`` -> burn:interval:steps;
    ``build_schedule(steps, burn, block::Vector)`` -> the block pattern repeated every ``block[end]`` steps
    (src/ParticlesMC.jl:255-261; the function itself is Arianna's)."""
    if isinstance(block, float):
        # build_schedule(interval, 0, base::Float64): logarithmic block 0, base^0, base^1, ... <= interval
        # (src/ParticlesMC.jl:254-256 calls it with base 2.0 for `log_base` schedulers)
        base, out, n = float(block), [burn], 0
        if base <= 1.0:
            raise NotImplementedError("logarithmic schedules need a base > 1")
        while burn + int(np.floor(base ** n)) <= steps:
            out.append(burn + int(np.floor(base ** n)))
            n += 1
        return sorted(set(out))
    if np.isscalar(block):
        return list(range(burn, steps + 1, int(block)))
    block = [int(b) for b in block]
    out, base = [], burn
    while base <= steps:
        out.extend(t + base for t in block if t + base <= steps)
        base += block[-1]
    return sorted(set(out))


class Simulation:
    def __init__(self, chains: Sequence[Particles], algorithm_list: Sequence[Dict], steps: int, *, path: str = "data",
                 verbose: bool = False, device: int = 0, chain_offset: int = 0, threads: int = 0,
                 mode: Optional[int] = None):
        self.chains = list(chains)
        self.algorithm_list = list(algorithm_list)
        self.steps = int(steps)
        self.path = path
        self.verbose = verbose
        self.t = 0
        metro = [a for a in self.algorithm_list if a["algorithm"] is Metropolis]
        if len(metro) != 1:
            raise ValueError("algorithm_list needs exactly one Metropolis entry")
        self.metropolis = metro[0]
        self.pool: Sequence[Move] = self.metropolis["pool"]
        self.sweepstep = int(self.metropolis.get("sweepstep", self.chains[0].N))
        self.ctx = make_context(self.chains, device=device, chain_offset=chain_offset, threads=threads, mode=mode)
        try:
            self.ctx.init_energy()
        except L.PMCError as e:
            if e.code == L.PMC_ERR_NONFINITE:
                raise ValueError("Initial configuration has infinite or NaN energy.") from e
            raise
        self.ctx.set_moves(pool_to_specs(self.pool))
        self.ctx.seed(int(self.metropolis.get("seed", 0)))
        self._sync_energy()

    # ---- host mirrors -------------------------------------------------------------------------------
    def _sync_energy(self):
        for s, e in zip(self.chains, self.ctx.energy()):
            s.energy[0] = e

    def _sync_state(self):
        pos, sp = self.ctx.download()
        for k, s in enumerate(self.chains):
            s.position[...] = pos[k]
            s.species[...] = sp[k]

    def _sync_counters(self):
        calls, acc = self.ctx.counters()
        for m, mv in enumerate(self.pool):
            mv.total_calls = int(calls[:, m].sum())
            mv.accepted_calls = int(acc[:, m].sum())
            mv._per_chain = (calls[:, m].copy(), acc[:, m].copy())

    def close(self):
        self.ctx.close()


def _append(path: str, line: str):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "a") as f:
        f.write(line)


def _write_frame(path: str, s: Particles, t: int, mode: str, fmt, last: bool):
    """One frame in the reference's grammar (``io.store_trajectory`` / ``io.store_lastframe``), 6 decimals."""
    from . import io as IO
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, mode) as f:
        (IO.store_lastframe if last else IO.store_trajectory)(f, s, t, fmt)


def _output_format(entry: Dict):
    """``fmt`` of a Store* entry: a Format instance or its name ("XYZ", "EXYZ", "LAMMPS"); default XYZ
    (src/ParticlesMC.jl:277-291)."""
    from . import io as IO
    f = entry.get("fmt", "XYZ")
    return getattr(IO, f)() if isinstance(f, str) else f


def run(sim: Simulation):
    """``run!(simulation)``: ``steps`` Arianna steps of ``sweepstep`` trials per chain."""
    outputs = [a for a in sim.algorithm_list if a["algorithm"] is not Metropolis]
    sched = [set(a.get("scheduler", [])) for a in outputs]
    fire_times = sorted(set().union(*sched)) if sched else []

    def fire(t: int):
        need_state = need_cnt = need_e = False
        for a, sc in zip(outputs, sched):
            if t not in sc:
                continue
            alg = a["algorithm"]
            need_e |= alg is StoreCallbacks
            need_cnt |= alg is StoreAcceptance
            need_state |= alg in (StoreTrajectories, StoreLastFrames)
        if need_e:
            sim._sync_energy()
        if need_cnt:
            sim._sync_counters()
        if need_state:
            sim._sync_state()
        for a, sc in zip(outputs, sched):
            if t not in sc:
                continue
            alg = a["algorithm"]
            if alg is StoreCallbacks:
                for cb in a.get("callbacks", ()):
                    for k, s in enumerate(sim.chains):
                        _append(os.path.join(sim.path, "chains", str(k + 1), f"{cb.__name__}.dat"), f"{t} {cb(s)!r}\n")
            elif alg is StoreAcceptance:
                for m, mv in enumerate(sim.pool):
                    rate = mv.accepted_calls / mv.total_calls if mv.total_calls else 0.0
                    _append(os.path.join(sim.path, "moves", str(m + 1), "acceptance.dat"), f"{t} {rate!r}\n")
            elif alg is StoreTrajectories:
                fmt = _output_format(a)
                for k, s in enumerate(sim.chains):
                    _write_frame(os.path.join(sim.path, "chains", str(k + 1), "trajectory" + fmt.extension), s, t, "a", fmt, False)
            elif alg is StoreLastFrames:
                fmt = _output_format(a)
                for k, s in enumerate(sim.chains):
                    _write_frame(os.path.join(sim.path, "chains", str(k + 1), "lastframe" + fmt.extension), s, t, "w", fmt, True)
            elif alg is PrintTimeSteps and sim.verbose:
                print(f"t = {t}")

    if 0 in fire_times and sim.t == 0:
        fire(0)
    while sim.t < sim.steps:
        nxt = min([t for t in fire_times if t > sim.t] + [sim.steps])
        sim.ctx.run((nxt - sim.t) * sim.sweepstep, sync=False)
        sim.t = nxt
        if nxt in fire_times:
            fire(nxt)
    sim.ctx.sync()
    sim._sync_energy()
    sim._sync_state()
    sim._sync_counters()
    return sim
