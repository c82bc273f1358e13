#!/usr/bin/env python
"""bench.py -- attempted Metropolis moves per second on the BASELINE.json workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|chains|box] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

One "step" = one pass of the hot path over the whole batch: every chain advanced by --sweeps sweeps
(1 sweep = N_particles attempted moves) in one pmc_run call.  Prints ONE JSON line (rank 0):

  value / e2e / roofline / clocks   BASELINE config 2: KA N=1000 x 4096 independent chains PER GPU (weak scaling)
  strong                            the same 4096 chains in TOTAL, split over the N GPUs (strong scaling)
  box                               BASELINE config 3: ONE KA box of N=2^20 particles, checkerboard sweeps, the cell
                                    grid cut into N slabs with NVLink halo pushes (strong scaling)
  cpu_baseline                      the oracle (C restatement of the reference CPU algorithm, one chain per host thread)
                                    on this box's cores, bounded sample, N = 1 only

  value      device-resident throughput: state already in HBM, K steps bracketed by CUDA events on the
             launch stream (max over ranks), L2 flushed between steps.
  e2e        the same metric through the C ABI with HOST buffers: every step uploads all states from pinned host
             memory (pmc_upload + pmc_init_energy), runs the sweeps and downloads energies and full states
             (pmc_energy + pmc_download).  The chains are held by --e2e-contexts library contexts on separate
             streams, staggered so that one context's copies overlap the others' sweeps (same chains, same bytes,
             every step's upload and download inside the timed region; each download feeds the next upload).
  roofline   pair-evaluation roofline of the sweep kernel (FP64 CUDA-core pipe):
               frac         reference-equivalent: moves/s x P x F with P = candidate pairs the REFERENCE evaluates per
                            move and F flops per pair, over the DFMA burst peak measured in this run
               frac_actual  the fp64 work the kernel really does: candidates that pass the integer prefilter, counted
                            on the device in this run (pmc_work_counters), x the fp64 flops per survivor of the SASS
               issue_frac   moves/s x warp-instructions per move (committed ncu capture) over the issue slots of the
                            GPU at the clock sampled in this run
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "attempted MC moves/sec"
UNIT = "moves/s"

# algorithmic work per attempted Displacement, reference-equivalent
WORK = {
    "chains": dict(P=2000.0, F=21.0 + 9.0 * 0.079),   # KA N=1000: 3x3x3 cells = whole box, old + new position
    "box": dict(P=1032.0, F=21.0 + 9.0 * 0.152),      # KA N=2^20: 27-cell stencil ~516 candidates, x2
}
# fp64 flops the kernels spend on ONE surviving candidate (old and new position together), counted in the SASS of the
# survivor loops (DFMA = 2): chains 14 DADD + 18 DFMA + 8 DMUL + 8 DSETP (minimum image per axis), box 8 DADD + 16 DFMA
# + 10 DMUL + 2 DSETP (cell frame, no minimum image).  DESIGN.md section 4.
FLOP_PER_SURVIVOR = {"chains": 14 + 2 * 18 + 8 + 8, "box": 8 + 2 * 16 + 10 + 2}
NCU_INDEX = os.path.join(ROOT, "profiles", "ncu_index.json")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "chains", "box"])
    ap.add_argument("--chains", type=int, default=4096, help="chains PER GPU of the weak-scaling leg; the strong leg splits this many over all GPUs")
    ap.add_argument("--particles", type=int, default=0, help="particles per chain (default 1000)")
    ap.add_argument("--box-particles", type=int, default=1 << 20)
    ap.add_argument("--sweeps", type=int, default=0, help="sweeps per step (default 10 chains / 8 box)")
    ap.add_argument("--equil", type=int, default=100, help="untimed equilibration sweeps from the lattice")
    ap.add_argument("--threads", type=int, default=0, help="CTA size of the sweep kernel (0 = library default)")
    ap.add_argument("--temperature", type=float, default=1.0)
    ap.add_argument("--precision", default="fp64", choices=["fp64", "mixed"],
                    help="mixed = fp32 pair terms / fp64 accumulation (reported separately, 1e-6 tolerance)")
    ap.add_argument("--prefilter", type=int, default=0, help="0 = fixed-point prefilter (default), -1 = visit all candidates in fp64")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--cpu-sweeps", type=int, default=600)
    ap.add_argument("--e2e-contexts", type=int, default=8,
                    help="chains workload, e2e leg: the chains are held by this many library contexts on separate streams, so "
                         "that one context's host<->device copies overlap another's sweeps (1 = a single context)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def workload_name(args, kind):
    if kind == "chains":
        N = args.particles or 1000
        return f"KA LJ 80:20 N={N} rho=1.2 T={args.temperature} x {args.chains} independent chains per GPU, Displacement sigma=0.05"
    return f"single KA LJ 80:20 box N={args.box_particles} rho=1.2 T={args.temperature}, checkerboard cell sweeps, Displacement sigma=0.05"


def config_dict(args):
    """The SAME dictionary for both arms (--impl ours / reference): it names the workload, nothing arm-specific."""
    kinds = ["chains", "box"] if args.workload == "all" else [args.workload]
    cfg = {"workload": workload_name(args, kinds[0]), "particles_per_chain": args.particles or 1000,
           "chains_per_gpu": args.chains, "sweeps_per_step": args.sweeps or 10, "temperature": args.temperature,
           "equilibration_sweeps": args.equil, "seed": 42, "precision": args.precision,
           "l2": "256 MiB buffer zeroed between steps (L2 flush)"}
    if "box" in kinds and len(kinds) > 1:
        cfg["also"] = workload_name(args, "box")
    return cfg


def ncu_figures(kind):
    """Figures of ONE launch of the dominant kernel from the committed `ncu --set full` summary (profiles/ncu_index.json
    names the CSV, the kernel it must hold and the moves of that launch).  None if the record is missing or stale."""
    try:
        ent = json.load(open(NCU_INDEX))[kind]
        rows = list(csv.DictReader(open(os.path.join(ROOT, ent["csv"]))))
        val = {r["metric"]: r["value"] for r in rows if r["launch"] == "0"}
        if ent["kernel"] not in val["Kernel Name"]:
            return None
        f = lambda k: float(val[k].replace(",", ""))
        return {"source": ent["csv"], "kernel": val["Kernel Name"], "moves_per_launch": ent["moves_per_launch"],
                "warp_instructions_per_move": f("smsp__inst_executed.sum") / ent["moves_per_launch"],
                "dram_read_bytes": f("dram__bytes_read.sum") * 1e6, "dram_write_bytes": f("dram__bytes_write.sum") * 1e6,
                "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "registers": int(f("launch__registers_per_thread")), "duration_ms": f("gpu__time_duration.sum")}
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def cpu_reference_arm(args, N, pos, sp, box, cores=None, one_chain=False):
    """The oracle on the host cores: one independent chain per thread (the reference's `parallel=true`)."""
    from oracle import oracle as O
    from particlesmc_b200 import models as M
    cores = cores or os.cpu_count() or 1
    par = M.flatten_model_matrix(M.KobAndersen())
    pool = O.make_pool([dict(kind="displacement", prob=1.0, sigma=0.05)])
    n = 1 if one_chain else cores
    systems = [O.OracleSystem(pos, sp, box, args.temperature, M.MODEL_LJ, par, O.LINKEDLIST) for _ in range(n)]
    return O, systems, pool


def time_cpu(O, systems, pool, n_trials, t0):
    t = time.perf_counter()
    # explicit thread count: torchrun exports OMP_NUM_THREADS=1, the reference arm must use all host cores
    used = O.run_chains(systems, 42, t0, n_trials, pool, revert_mode=0, n_threads=min(len(systems), os.cpu_count() or 1))
    dt = time.perf_counter() - t
    return len(systems) * n_trials / dt, dt, min(used, len(systems))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (its C restatement: Julia is not available on this
    box) on all host cores, on the chains workload of the `ours` arm; each step is a bounded sample of it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from particlesmc_b200.synthetic import ka_lattice
    N = args.particles or 1000
    sweeps = args.sweeps or 10
    pos, sp, box = ka_lattice(N, 1.2, seed=0)
    ref_sweeps = max(1, min(sweeps, 4))
    O, systems, pool = cpu_reference_arm(args, N, pos, sp, box)
    t0 = 0
    for _ in range(args.warmup):
        time_cpu(O, systems, pool, max(1, N // 10), t0)
        t0 += max(1, N // 10)
    tot_t, tot_moves, cores = 0.0, 0, 1
    for _ in range(args.steps):
        v, dt, cores = time_cpu(O, systems, pool, ref_sweeps * N, t0)
        t0 += ref_sweeps * N
        tot_t += dt
        tot_moves += len(systems) * ref_sweeps * N
    value = tot_moves / tot_t
    sample = (f"{len(systems)} chains (one per host thread) x {ref_sweeps} sweeps of N={N} per step: a bounded sample of the "
              f"{args.chains} x {sweeps}-sweep step of the GPU arm (chains are independent, throughput does not depend on their number)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args),
            "note": "C restatement of the reference CPU algorithm (oracle/), not Julia: Julia and Arianna.jl are not available on this box",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
class Env:
    pass


def measure(env, args, kind, Mc, strong_chains=False):
    """One workload on this rank's GPU: device-resident throughput, e2e through host buffers, work counters.
    kind = 'chains' (Mc chains on this rank) or 'box' (one box over all ranks)."""
    import torch
    import torch.distributed as dist
    from particlesmc_b200 import _lib as L
    from particlesmc_b200 import models as M
    from particlesmc_b200.device import DeviceContext
    from particlesmc_b200.synthetic import ka_lattice

    rank, world, local, dev = env.rank, env.world, env.local, env.dev
    box_mode = kind == "box"
    N = args.box_particles if box_mode else (args.particles or 1000)
    sweeps = args.sweeps or (8 if box_mode else 10)
    pos, sp, box = ka_lattice(N, 1.2, seed=0)
    par = M.flatten_model_matrix(M.KobAndersen())
    mode = L.MODE_BOX if box_mode else L.MODE_CHAINS
    prec = L.MIXED if (args.precision == "mixed" and not box_mode) else L.FP64
    # chains are keyed by their GLOBAL index: rank r holds chains [r*Mc, (r+1)*Mc)
    ctx = DeviceContext(Mc, N, 3, 2, M.MODEL_LJ, mode=mode, device=local, chain_offset=rank * Mc, threads=args.threads,
                        prefilter=args.prefilter, precision=prec)
    ctx.set_stream(env.stream.cuda_stream)
    ctx.set_model(par)

    # pinned host buffers in the caller's (reference) layout
    h_pos = torch.empty((Mc, N, 3), dtype=torch.float64).pin_memory()
    h_sp = torch.empty((Mc, N), dtype=torch.int64).pin_memory()
    h_box = torch.empty((Mc, 3), dtype=torch.float64).pin_memory()
    h_T = torch.full((Mc,), args.temperature, dtype=torch.float64).pin_memory()
    h_E = torch.empty((Mc,), dtype=torch.float64).pin_memory()
    h_pos.numpy()[:] = pos[None]
    h_sp.numpy()[:] = sp[None]
    h_box.numpy()[:] = box[None]

    def upload():
        ctx.upload_raw(h_pos.data_ptr(), h_sp.data_ptr(), h_box.data_ptr(), h_T.data_ptr(), 0, Mc)
        ctx.init_energy()

    def download():
        ctx.energy_into(h_E.data_ptr())
        ctx.download_raw(h_pos.data_ptr(), h_sp.data_ptr(), 0, Mc)

    upload()
    split = box_mode and world > 1
    if split:  # one box, its cell grid cut into one slab per rank; moves pushed over NVLink peer memory
        from particlesmc_b200.sharding import attach_box_peers
        attach_box_peers(ctx)
    ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
    ctx.seed(42)
    trials_per_step = sweeps * N
    if args.equil > 0:
        ctx.run(args.equil * N)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -----------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        env.flush.zero_()
        ctx.run(trials_per_step, sync=False)
    barrier()
    # nvidia-smi needs ~0.2 s to deliver its first sample: keep the GPU under the same load (untimed steps, all
    # ranks alike) until it is up, so that the clock record covers the timed region even when a step is short.
    # Every rank must run the same number of sweeps of the split box, so the count comes from an all-reduced step time.
    t_s = time.perf_counter()
    ctx.run(trials_per_step, sync=True)
    tt = torch.tensor([time.perf_counter() - t_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    n_load = int(min(400, np.ceil(0.35 / max(float(tt.item()), 1e-4))))
    for _ in range(n_load):
        ctx.run(trials_per_step, sync=False)
    ctx.sync()
    launches0 = ctx.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        env.flush.zero_()
        ctx.run(trials_per_step, sync=False)
    ev[1].record()
    barrier()
    ms_total = ev[0].elapsed_time(ev[1])
    launches = ctx.launch_count() - launches0
    for _ in range(max(1, n_load // 2)):  # keep the same load up until the sampler has seen it (same count on every rank)
        ctx.run(trials_per_step, sync=False)
    ctx.sync()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["note"] = "sampled every 50 ms from the last warm-up step through the timed region (same load)"
    # per-launch duration of the sweep kernel(s), CUDA events recorded inside the library on the same stream
    kernel_ms = []
    for _ in range(args.steps):
        ctx.run(trials_per_step, sync=True)
        kernel_ms.append(ctx.last_run_ms())
    # work really done, counted on the device during one more (untimed) step
    ctx.work_counters(1)
    ctx.run(trials_per_step, sync=True)
    survivors, evaluations = ctx.work_counters(0)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    moves_local = Mc * trials_per_step // (world if split else 1)   # moves this rank's GPU processes per step
    moves_per_step = Mc * trials_per_step * (1 if split else world)  # whole job
    value = moves_per_step * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI with host buffers -------------------------------------------------
    e2e = None
    if not args.no_e2e:
        K = 1
        if not box_mode:
            K = max(1, args.e2e_contexts)
            while K > 1 and (Mc % K != 0 or Mc // K < 128):  # at least 128 chains per context
                K //= 2
        parts = []
        if K > 1:
            # the same chains (same global indices, same seed) split over K contexts with their own streams: while one
            # context sweeps, the next one's upload and the previous one's download use the copy engines
            Mk = Mc // K
            for k in range(K):
                c = DeviceContext(Mk, N, 3, 2, M.MODEL_LJ, mode=mode, device=local, chain_offset=rank * Mc + k * Mk,
                                  threads=args.threads, prefilter=args.prefilter, precision=prec)
                st = torch.cuda.Stream(device=dev)
                c.set_stream(st.cuda_stream)
                c.set_model(par)
                parts.append((c, st, k * Mk))

            def up(c, o):
                c.upload_raw(h_pos.data_ptr() + o * N * 3 * 8, h_sp.data_ptr() + o * N * 8, h_box.data_ptr() + o * 3 * 8,
                             h_T.data_ptr() + o * 8, 0, Mk)
                c.init_energy()

            def down(c, o):
                c.energy_into(h_E.data_ptr() + o * 8)
                c.download_raw(h_pos.data_ptr() + o * N * 3 * 8, h_sp.data_ptr() + o * N * 8, 0, Mk)
            for c, st, o in parts:
                up(c, o)
                c.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
                c.seed(42)

            # Every step, every context: upload its chains, sweep, download them (the next step's upload reads what
            # this step's download wrote: the state round-trips through the host).  The contexts are staggered: while
            # one is copying, the others sweep -- context k's download of step s is followed at once by its upload and
            # launch of step s + 1, so the pipeline fills and drains once per e2e_steps() call, not once per step.
            def e2e_steps(n):
                for s in range(n):
                    for c, st, o in parts:
                        if s > 0:
                            down(c, o)
                        up(c, o)
                        c.run(trials_per_step, sync=False)
                for c, st, o in parts:
                    down(c, o)
        else:
            def e2e_steps(n):
                for _ in range(n):
                    upload()
                    ctx.run(trials_per_step, sync=False)
                    download()
        e2e_steps(2)
        barrier()
        t0 = time.perf_counter()
        ev[0].record()
        e2e_steps(args.steps)
        ev[1].record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e_ms = max(ev[0].elapsed_time(ev[1]), wall_ms)
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d = h_pos.numel() * 8 + h_sp.numel() * 8 + h_box.numel() * 8 + h_T.numel() * 8
        d2h = h_pos.numel() * 8 + h_sp.numel() * 8 + h_E.numel() * 8
        for c, st, o in parts:
            c.close()
        e2e = {"value": moves_per_step * args.steps / (float(t.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "contexts": K,
               "energy_per_particle_mean": float(h_E.numpy().mean() / N)}

    # ---- energy sanity in the same run: bookkeeping vs recomputed --------------------------------------
    e_run, e_tot = ctx.energy(), ctx.total_energy()
    drift = float(np.max(np.abs(e_run - e_tot) / np.abs(e_tot)))
    calls, acc = ctx.counters()
    ctx.close()
    return dict(kind=kind, N=N, Mc=Mc, sweeps=sweeps, value=value, ms_total=ms_total, launches=int(launches), clocks=clocks,
                kernel_ms=float(np.mean(kernel_ms)), e2e=e2e, survivors_per_move=survivors / max(moves_local, 1),
                evaluations_per_move=evaluations / max(moves_local, 1), moves_local=moves_local, split=split,
                checks={"energy_bookkeeping_rel_drift": drift, "acceptance": float(acc.sum() / max(calls.sum(), 1)),
                        "energy_per_particle": float(e_tot.mean() / N)})


def roofline(env, args, m, peak64, peak32):
    kind = m["kind"]
    work = WORK[kind]
    mixed = args.precision == "mixed" and kind == "chains"
    peak = peak32 if mixed else peak64
    per_gpu = m["moves_local"] / (m["kernel_ms"] * 1e-3)  # moves/s of this GPU while the kernel runs
    achieved = per_gpu * work["P"] * work["F"] / 1e12
    actual = per_gpu * m["survivors_per_move"] * FLOP_PER_SURVIVOR[kind] / 1e12
    ncu = ncu_figures(kind) if (args.precision == "fp64" and args.prefilter == 0) else None
    clk = (m["clocks"] or {}).get("sm_mhz") or 1965.0
    issue_frac = None
    if ncu:
        issue_frac = per_gpu * ncu["warp_instructions_per_move"] / (env.sms * 4 * clk * 1e6)
    if kind == "chains":
        kernel = ("k_chain_sweep_spec<MIXED>" if mixed else "k_chain_sweep_spec") if args.prefilter == 0 else \
            ("k_chain_sweep_fast" if args.prefilter == 1 else "k_chain_sweep")
    else:
        kernel = "k_box_sweep_all (8 colours in one launch) + 4 cell-rebuild kernels + reduction"
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return {"bound": "fp32_pipe" if mixed else "fp64_pipe", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "frac_actual": actual / peak, "issue_frac": issue_frac,
            "traffic": (ncu["dram_read_bytes"] + ncu["dram_write_bytes"]) if ncu else None, "kernel": kernel,
            "kernel_ms_per_launch": m["kernel_ms"], "launches_per_step": m["launches"] / args.steps,
            "survivors_per_move": m["survivors_per_move"], "evaluations_per_move": m["evaluations_per_move"],
            "fp64_flop_per_survivor": FLOP_PER_SURVIVOR[kind],
            "work_model": f"frac: reference-equivalent, P={work['P']:.0f} candidate pairs/move x F={work['F']:.2f} flop/pair "
                          "(SURVEY.md 8d); frac_actual: candidates that passed the integer prefilter (counted on the device in "
                          "this run) x fp64 flops per survivor of the kernel's SASS; issue_frac: warp-instructions per move of "
                          "the committed ncu capture x moves/s over SMs x 4 schedulers x sampled clock",
            "note": "the kernel rejects ~91 % of the reference's candidates with a 3-instruction integer test and evaluates only the "
                    "survivors in fp64, so `frac` is not a pipe utilisation -- `frac_actual` and `issue_frac` are",
            "ncu": ncu,
            "peak_source": "FMA burst micro-benchmark run in this process (pmc_measure_fma_peak); MEASURED_PEAKS.json holds no "
                           "FP64/FP32 CUDA-core figure",
            "fp64_fma_peak_tflops": peak64, "fp32_fma_peak_tflops": peak32,
            "hbm_algorithmic_gbs": 2.0 * m["Mc"] * m["N"] * 28 / (m["kernel_ms"] * 1e-3) / 1e9 / (env.world if m["split"] else 1),
            "hbm_peak_gbs": json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from particlesmc_b200.device import measure_fma_peak

    env = Env()
    env.rank = int(os.environ.get("RANK", "0"))
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    env.local = int(os.environ.get("LOCAL_RANK", "0"))
    if env.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", env.local))
    torch.cuda.set_device(env.local)
    env.dev = torch.device("cuda", env.local)
    env.sms = torch.cuda.get_device_properties(env.local).multi_processor_count
    env.stream = torch.cuda.Stream(device=env.dev)  # a non-default stream shared by torch events and the library
    torch.cuda.set_stream(env.stream)
    env.flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.dev)  # > 126 MB L2
    rank, world = env.rank, env.world

    kinds = ["chains", "box"] if args.workload == "all" else [args.workload]
    res = {}
    if "chains" in kinds:
        res["chains"] = measure(env, args, "chains", args.chains)
        if world > 1 and not args.no_strong and args.chains % world == 0:
            res["strong"] = measure(env, args, "chains", args.chains // world, strong_chains=True)
    if "box" in kinds:
        res["box"] = measure(env, args, "box", 1)

    if rank == 0:
        peak64 = measure_fma_peak(True, env.local)
        peak32 = measure_fma_peak(False, env.local)
        head = res[kinds[0]]
        cpu = None
        if not args.no_cpu_baseline and world == 1 and "chains" in res:
            from particlesmc_b200.synthetic import ka_lattice
            N = res["chains"]["N"]
            pos, sp, box = ka_lattice(N, 1.2, seed=0)
            O, systems, pool = cpu_reference_arm(args, N, pos, sp, box)
            time_cpu(O, systems, pool, N // 4, 0)
            v, dt, cores = time_cpu(O, systems, pool, args.cpu_sweeps * N, N)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
                   "sample": f"{len(systems)} chains (one per host thread) x {args.cpu_sweeps} sweeps of N={N}"}

        def sub(m, scaling):
            return {"value": m["value"], "unit": UNIT, "scaling": scaling, "ms_per_step": m["ms_total"] / args.steps,
                    "sweeps_per_step": m["sweeps"], "ms_per_sweep": m["ms_total"] / args.steps / m["sweeps"],
                    "systems_per_gpu": m["Mc"], "particles": m["N"], "e2e": m["e2e"], "gpu_launches": m["launches"],
                    "clocks": m["clocks"], "roofline": roofline(env, args, m, peak64, peak32), "checks": m["checks"]}

        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": head["ms_total"] / args.steps, "higher_is_better": True,
                "scaling": "strong" if head["split"] else "weak", "vs_baseline": None,
                "dtype": "f64" if args.precision == "fp64" else "f32 pair terms / f64 accumulate", "data": "synthetic",
                "config": config_dict(args), "clocks": head["clocks"], "e2e": head["e2e"],
                "gpu_launches": sum(m["launches"] for m in res.values()),
                "roofline": roofline(env, args, head, peak64, peak32), "cpu_baseline": cpu, "checks": head["checks"]}
        if "strong" in res:
            line["strong"] = sub(res["strong"], "strong")
            line["strong"]["note"] = f"the same {args.chains} chains in total, {args.chains // world} per GPU"
        elif "chains" in res and world == 1:
            line["strong"] = {"value": res["chains"]["value"], "unit": UNIT, "scaling": "strong",
                              "note": "N = 1: identical to the headline run"}
        if "box" in res and kinds[0] != "box":
            line["box"] = sub(res["box"], "strong")
            line["box"]["workload"] = workload_name(args, "box")
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
