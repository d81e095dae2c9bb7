#!/usr/bin/env python
"""bench.py — monomer MC updates/s of the fixed-force MCMC hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" is one pass of the hot path over one batch: S trials (default 500 = one --stepout
interval) on every chain of the workload.  Workload C2 (BASELINE.json configs[1]): interacting
dielectric chains, n=512, 4096 independent replicas PER GPU (weak scaling), E0=1, K1=1, K2=0, kT=1,
b=1, Fz=0.5, default step sizes with adaptation on, stepout=500.

One JSON line is printed by rank 0.  Under torchrun each rank drives one GPU; chains are sharded
by global chain id with no data-path collective; the only collective is the final gather of the
per-chain averages (NCCL), which is inside the e2e timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "polymer-stats_b200"))

METRIC = "monomer MC updates/sec (whole box), interacting n=512 dielectric chains"
UNIT = "updates/s"

WORKLOADS = {
    # name: (case kwargs, chains per GPU, trials per step, stepout)
    "C2": (dict(n=512, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.5, Fx=0.0, chain_type="dielectric",
                energy_type="interacting"), 4096, 500, 500),
    "C3": (dict(n=512, E0=1.0, mu=0.5, kT=1.0, b=1.0, Fz=1.0, Fx=0.0, chain_type="polar",
                energy_type="interacting"), 4096, 500, 500),
    "C4": (dict(n=100, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.5, Fx=0.0, chain_type="dielectric",
                energy_type="noninteracting"), 16384, 20000, 500),
    "C5": (dict(n=4096, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.5, Fx=0.0, chain_type="dielectric",
                energy_type="interacting"), 148, 200, 200),  # ≥ 200 trials per chain (one wave of CTAs:
    # the launch ends with its slowest chain, and the spread of Σ idx(n−1−idx) shrinks as 1/sqrt(trials))
    # the clustering driver (mcmc_clustering_eap_chain.jl), shapes of its launchers:
    # run/phases-kT-small-n_2023-09-09.jl (all-pairs, bending), run/Ising_2025-12-17.jl, run/phases-big_2023-05-18.jl
    "K1": (dict(n=100, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.25, Fx=0.0, chain_type="dielectric",
                energy_type="interacting", kappa=0.5, clustering=True, adj_ub=0.40), 4096, 500, 250),
    "K2": (dict(n=100, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.25, Fx=0.0, chain_type="dielectric",
                energy_type="Ising", kappa=0.5, clustering=True, adj_ub=0.40), 65536, 2000, 250),
    "K3": (dict(n=400, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.25, Fx=0.0, chain_type="dielectric",
                energy_type="cutoff", cutoff_radius=7.5, kappa=0.5, clustering=True, adj_ub=0.40), 2368, 200, 100),
    # the size of the reference's own studies: 20 cases × 25 runs (run/Ising_2025-12-17.jl) — one chain per warp /
    # more warps per chain (DESIGN.md §3.1 "Small ensembles", §9)
    "K4": (dict(n=100, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.25, Fx=0.0, chain_type="dielectric",
                energy_type="Ising", kappa=0.5, clustering=True, adj_ub=0.40), 500, 20000, 2500),
    "K5": (dict(n=100, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.25, Fx=0.0, chain_type="dielectric",
                energy_type="interacting", kappa=0.5, clustering=True, adj_ub=0.40), 500, 2000, 250),
}
SEED = 20260101


def flops_per_update(n: int, energy_type: str) -> float:
    """Algorithmic FP64 flop per update (SURVEY.md §8d / BASELINE.md §3)."""
    if energy_type in ("interacting", "cutoff"):
        return 2.0 * 34.0 * ((n - 1) * (n - 2) / 6.0 + (n - 1))
    return 210.0 if energy_type == "Ising" else 66.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smmax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def metric_of(workload, kw):
    if workload == "C2":
        return METRIC
    drv = "clustering driver, " if kw.get("clustering") else ""
    return f"monomer MC updates/sec (whole box), {drv}{kw['energy_type']} n={kw['n']} {kw['chain_type']} chains"


def make_config(workload, kw, per_gpu, S, stepout):
    """The `config` object — identical for both arms."""
    drv = "clustering driver (move! + cluster_flip!), " if kw.get("clustering") else ""
    return {"workload": f"{workload}: {drv}{kw['energy_type']} {kw['chain_type']} chains n={kw['n']}, "
                        f"{per_gpu} replicas per GPU, {S} trials per step, stepout={stepout}",
            "chains_per_gpu": per_gpu, "trials_per_step": S, "seed": SEED,
            "l2": "flushed between timed steps (256 MiB write)", **kw}


def cpu_reference_run(kw, n_chains, trials, threads, algo=0):
    """The reference algorithm (deep-copy-free full energy recompute per trial, oracle algo 0) on
    host threads, one chain per thread at a time."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    oc = O.make_case(**kw)
    secs = O.bench(oc, SEED, algo, n_chains, trials, threads)
    return n_chains * trials / secs, secs


def sized_cpu_sample(kw, threads, target_s, algo=0):
    """Pick trials/chain so that `threads` chains take about target_s seconds."""
    probe_trials = 20 if kw["energy_type"] == "interacting" and kw["n"] >= 512 else 2000
    if kw["n"] >= 4096:
        probe_trials = 2
    ups, secs = cpu_reference_run(kw, threads, probe_trials, threads, algo)
    per_thread_rate = ups / threads
    trials = max(probe_trials, int(per_thread_rate * target_s))
    return trials


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's own CPU algorithm for the path on the box's host cores.
    Julia is not installable here, so the timed code is the oracle's C restatement of the reference
    algorithm (kind "port"): full O(n²) energy recompute per trial, one chain per thread."""
    if rank != 0:
        return 0
    kw, per_gpu, S, stepout = WORKLOADS[args.workload]
    if args.chains_per_gpu:
        per_gpu = args.chains_per_gpu
    if args.trials_per_step:
        S = args.trials_per_step
        stepout = min(stepout, S)
    threads = os.cpu_count() or 1
    trials = sized_cpu_sample(kw, threads, target_s=args.ref_step_seconds)
    for _ in range(args.warmup):
        cpu_reference_run(kw, threads, max(1, trials // 10), threads)
    t_total, updates = 0.0, 0
    for _ in range(args.steps):
        ups, secs = cpu_reference_run(kw, threads, trials, threads)
        t_total += secs
        updates += threads * trials
    value = updates / t_total
    sample = (f"{threads} chains x {trials} trials per step (one chain per thread), n={kw['n']} "
              f"{kw['energy_type']} {kw['chain_type']}, same parameters as the GPU arm")
    line = {
        "impl": "reference", "metric": metric_of(args.workload, kw), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args.workload, kw, per_gpu, S, stepout),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference algorithm (oracle/polymc_oracle.c, algo 0); "
                                 "Julia is not installed in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """Exactly one JSON line on the real stdout (libraries such as NCCL may print to fd 1; everything
    else is redirected to stderr for the duration of the run)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--chains-per-gpu", type=int, default=None)
    ap.add_argument("--trials-per-step", type=int, default=None)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--ref-step-seconds", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import polymc as pm

    if not torch.cuda.is_available() or pm.device_count() < 1:
        emit({"error": "no CUDA device: libpolymc_b200 has no CPU fallback"})
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    kw, per_gpu, S, stepout = WORKLOADS[args.workload]
    if args.chains_per_gpu:
        per_gpu = args.chains_per_gpu
    if args.trials_per_step:
        S = args.trials_per_step
        stepout = min(stepout, S)
    n = kw["n"]
    case = pm.make_case(**kw)
    ens = pm.Ensemble(case, replicas=per_gpu, seed=SEED, device=local_rank, chain_id_base=rank * per_gpu)
    stream = torch.cuda.current_stream(dev)
    ens.set_stream(stream.cuda_stream)

    # L2 flush buffer (larger than the 126 MB L2), written between timed steps
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- measured FP64 roofline denominator (DFMA microbenchmark, before the clocks heat up) ----
    probe_tf = max(pm.fp64_peak_probe(local_rank, 1 << 16)[0] for _ in range(3))

    clustering = bool(kw.get("clustering"))
    if clustering:
        ens.begin_stage(1.0)   # a fresh mcmc(nsteps, pargs, chain) call (mcmc_clustering_eap_chain.jl:171-265)

    def run_resident():
        if clustering:
            ens.run_ex(S, stepout, fetch_rows=False)
        else:
            ens.run(S, stepout, fetch_rows=False)

    # ---- warm-up -------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 0)):
        run_resident()

    # ---- value: state resident in HBM, device-timed -----------------------------------------------
    sampler = ClockSampler(local_rank)
    launches0 = ens.launch_count()
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    step_ms, kernel_ms = [], []
    for _ in range(args.steps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run_resident()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        kernel_ms.append(ens.last_run_ms())
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = ens.launch_count() - launches0
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    updates_per_step = per_gpu * S * world
    value = updates_per_step * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers in, host buffers out, through the public Ensemble API -----------------
    e2e = None
    if not args.no_e2e:
        phi_h = torch.empty((per_gpu, n), dtype=torch.float64).pin_memory()
        th_h = torch.empty((per_gpu, n), dtype=torch.float64).pin_memory()
        p0, t0 = ens.get_state_all()
        phi_h.numpy()[:] = p0
        th_h.numpy()[:] = t0
        rows = max(1, S // stepout) if stepout > 0 else 0
        traj_h = torch.empty((per_gpu, rows, 8), dtype=torch.float64).pin_memory()
        roll_h = torch.empty((per_gpu, rows, 17), dtype=torch.float64).pin_memory()
        # keep the step counter aligned with stepout so that every e2e step emits `rows` rows
        def e2e_step():
            ens.set_state_all(phi_h.numpy(), th_h.numpy())                 # H2D + cache rebuild
            ens.run(S, stepout, traj=traj_h.numpy(), roll=roll_h.numpy())  # hot loop + D2H rows (17 columns)
            avg, ar, nrm = ens.averages()                                  # D2H results
            if world > 1:                                                  # final gather of averages (NCCL)
                t = torch.from_numpy(np.concatenate([avg, ar[:, None], nrm[:, None]], axis=1)).to(dev)
                outs = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(outs, t)
                torch.cuda.synchronize(dev)
            return avg
        e2e_step()
        barrier()
        t0w = time.perf_counter()
        for _ in range(args.steps):
            avg = e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0w], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = phi_h.numel() * 8 + th_h.numel() * 8
        d2h = traj_h.numel() * 8 + roll_h.numel() * 8 + per_gpu * (16 + 2) * 8
        e2e = {"value": updates_per_step * args.steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * float(dt.item()) / args.steps,
               "api": "polymc.Ensemble.set_state_all + run(host traj/roll) + averages (C ABI pmc_*), pinned host buffers"}

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    F = flops_per_update(n, kw["energy_type"])
    if clustering:
        # changed pair terms of a composite trial: segment×everything + heads×tails, from the measured cluster
        # sizes (the single-monomer formula with the segment in place of idx): informational
        cs = ens.cluster_stats()
        trials = ens.diagnostics()[:, 5].sum()
        mean_seg = 1.0 + (cs[:, 1].sum() - cs[:, 0].sum()) / max(1.0, trials)
        if kw["energy_type"] in ("interacting", "cutoff"):
            F = 2.0 * 34.0 * ((n - mean_seg) * (n - mean_seg - 1) / 6.0 + mean_seg * (n - 1))
    kavg_ms = sum(kernel_ms) / len(kernel_ms)
    achieved_tf = per_gpu * S * F / (kavg_ms * 1e-3) / 1e12
    derived_tf = 148 * 64 * 2 * 1.965e9 / 1e12
    kernel = "k_run_cta_win" if kw["energy_type"] == "interacting" and n <= 3000 else (
        "k_run_cta" if kw["energy_type"] == "interacting" else ("k_run_warp" if per_gpu < 20000 else "k_run_lane"))
    if clustering:  # the library picks the packing of the O(1)-energy kernels by chain count (polymc.cu)
        kernel = "k_run_cta_cluster" if kw["energy_type"] in ("interacting", "cutoff") else (
            "k_run_warp_cluster" if per_gpu < 11000 else "k_run_lane_cluster")
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the C2 launch, from
    # profiles/r01d_dram_traffic_k_run_cta_win_bench_launch.csv (ncu, same command): 105.3 MB read (the
    # chain records, once per launch) + 2.8-6.3 MB written.  Other workloads: not captured.
    traffic = 1.10e8 if (args.workload == "C2" and per_gpu == 4096 and S == 500) else None
    roofline = {
        "bound": "fp64", "kernel": kernel, "achieved": achieved_tf, "peak": probe_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / probe_tf, "traffic": traffic,
        "peak_source": "measured in this run: DFMA-only microbenchmark pmc_fp64_peak_probe (MEASURED_PEAKS.json "
                       "holds no FP64 figure)",
        "peak_derived": derived_tf, "frac_of_derived": achieved_tf / derived_tf,
        "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum); the path is FP64-pipe "
                        "bound, HBM traffic is ~0.02 % of peak",
        "flop_per_update": F, "updates_per_launch": per_gpu * S, "kernel_ms_avg": kavg_ms,
        "kernel_share_of_step": kavg_ms * len(kernel_ms) / sum(step_ms),
    }

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        trials = sized_cpu_sample(kw, threads, args.cpu_baseline_seconds)
        ups, secs = cpu_reference_run(kw, threads, trials, threads, algo=0)
        cpu_baseline = {"value": ups, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{threads} chains x {trials} trials (one chain per thread, {secs:.1f} s), same "
                                  f"parameters; oracle algo 0 = the reference's full-recompute algorithm"}

    if rank == 0:
        line = {
            "metric": metric_of(args.workload, kw), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args.workload, kw, per_gpu, S, stepout),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "wall_s_timed_region": t_wall,
        }
        emit(line)
    ens.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
