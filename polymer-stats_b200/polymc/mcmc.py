"""Host-side mirror of mcmc_eap_chain.jl: the ArgParse table (:19-153), `mcmc(nsteps, pargs)`
(:171-376) and the result printing (:386-395), driving libpolymc_b200.so through the C ABI.

The reference's toolchain (Julia) is not present in this image, so this Python module plays the
role of the Julia host; polymer-stats_b200/julia/mcmc_eap_chain.jl is the `ccall` twin.

Same option names, defaults, argument meaning and error behaviour as the reference.  Additive
options (not in the reference): --replicas, --seed, --device, --pair-precision.
"""
from __future__ import annotations

import argparse
import math
import sys
import time

import numpy as np

from . import lib
from .output import ROLL_HEADER, TRAJ_HEADER, result_lines, write_rows


def build_parser() -> argparse.ArgumentParser:
    """The @add_arg_table of mcmc_eap_chain.jl:19-153, option for option."""
    p = argparse.ArgumentParser(prog="mcmc_eap_chain", allow_abbrev=False,
                                description="fixed-force MCMC of an electro-active polymer chain (B200 path)")
    a = p.add_argument
    a("--E0", "-e", type=float, default=0.0, help="magnitude of the electric field")
    a("--chain-type", "-T", type=str, default="dielectric", help="chain type (dielectric|polar)")
    a("--K1", "-J", type=float, default=1.0, help="dipole susceptibility along the monomer axis")
    a("--K2", "-K", type=float, default=0.0, help="dipole susceptibility orthogonal to the monomer axis")
    a("--mu", "-m", type=float, default=1e-2, help="dipole magnitude (electret chain)")
    a("--energy-type", "-u", type=str, default="noninteracting",
      help="energy type (noninteracting|interacting|Ising)")
    a("--kT", "-k", type=float, default=1.0, help="dimensionless temperature")
    a("--ensemble-type", "-E", type=str, default="force", help="ensemble (force|end-to-end)")
    a("--Fz", "-F", type=float, default=0.0, help="force in the z-direction (direction of E-field)")
    a("--Fx", "-G", type=float, default=0.0, help="force in the x-direction")
    a("--rz", "-z", type=float, default=0.0, help="end-to-end vector z (end-to-end ensemble only)")
    a("--rx", "-x", type=float, default=0.0, help="end-to-end vector x (end-to-end ensemble only)")
    a("--mlen", "-b", type=float, default=1.0, help="monomer length")
    a("--num-monomers", "-n", type=int, default=100, help="number of monomers")
    a("--num-steps", "-N", type=int, default=100000, help="number of steps")
    a("--num-inits", "-M", type=int, default=1, help="number of random initializations (B200 path: the acceptor is re-bound to the new chain after a "
                                                          "re-initialisation; the reference leaves it stale, mcmc_eap_chain.jl:360 — see DESIGN.md §5)")
    a("--force-init", "-I", action="store_true", help="force (no acceptance test) every initialization")
    a("--phi-step", "-p", type=float, default=3 * math.pi / 8, help="maximum phi step length")
    a("--do-flips", action="store_true", help="trial moves with flipping monomers")
    a("--theta-step", "-q", type=float, default=3 * math.pi / 16, help="maximum theta step length")
    a("--chain-frac-step", "-f", type=float, default=0.15, help="(end-to-end ensemble only)")
    a("--step-adjust-lb", "-L", type=float, default=0.15, help="lower acceptance bound for step adaptation")
    a("--step-adjust-ub", "-U", type=float, default=0.55, help="upper acceptance bound for step adaptation")
    a("--step-adjust-scale", "-A", type=float, default=1.1, help="step adaptation factor (1.0 disables)")
    a("--steps-per-adjust", "-S", type=int, default=2500, help="steps between step-size adjustments")
    a("--acc", "-a", type=str, default="metropolis", help="acceptance function (metropolis)")
    a("--umbrella-sampling", "-B", action="store_true", help="use umbrella sampling")
    a("--update-freq", type=float, default=15.0, help="progress update frequency (seconds)")
    a("--verbose", "-v", type=int, default=3, help="verbosity 0-3")
    a("--prefix", "-P", type=str, default="eap-mcmc", help="prefix for output files")
    a("--postfix", "-Q", type=str, default="", help="postfix for output files (parsed, unused)")
    a("--stepout", "-s", type=int, default=500, help="steps between storing microstates")
    a("--numeric-type", type=str, default="float64", help="accumulator type (float64|float128|dec128|big)")
    a("--profile", "-Z", action="store_true", help="profile the program")
    # additive
    a("--replicas", type=int, default=1, help="[B200 path] independent replica chains run concurrently and pooled")
    a("--seed", type=int, default=None, help="[B200 path] Philox seed (default: time-based, like the unseeded reference)")
    a("--device", type=int, default=0, help="[B200 path] CUDA device index")
    a("--pair-precision", default="fp64", choices=["fp64", "fp32"],
      help="[B200 path] fp32: the rectangle of unchanged-dipole pairs of a trial in FP32 (tolerance in include/polymc.h); "
           "everything else FP64")
    return p


def parse_args(argv=None) -> dict:
    """Returns the reference's `pargs` Dict (keys are the long option names, e.g. "num-steps")."""
    ns = build_parser().parse_args(argv)
    return {k.replace("_", "-"): v for k, v in vars(ns).items()}


def default_pargs(**overrides) -> dict:
    d = parse_args([])
    for k, v in overrides.items():
        d[k.replace("_", "-")] = v
    return d


def _log(pargs, level, msg):
    """ConsoleLogger to stderr by --verbose (mcmc_eap_chain.jl:157-165): 3 info, 2 warn, 1 error."""
    need = {"info": 3, "warn": 2, "error": 1}[level]
    if pargs.get("verbose", 3) >= need:
        tag = {"info": "Info", "warn": "Warning", "error": "Error"}[level]
        print(f"[ {tag}: {msg}", file=sys.stderr)


def case_from_pargs(pargs: dict) -> lib.PmcCase:
    """EAPChain(pargs) argument mapping (inc/eap_chain.jl:60-135)."""
    if pargs["energy-type"] == "cutoff":
        # this driver has no --cutoff-radius, so the reference dies on pargs["cutoff-radius"]
        # (inc/eap_chain.jl:102); the cut-off energy belongs to mcmc_clustering_eap_chain
        raise lib.PolymcError(-1, "energy-type is not understood. ('cutoff' needs --cutoff-radius: use "
                                  "mcmc_clustering_eap_chain)")
    return lib.make_case(
        n=pargs["num-monomers"], E0=pargs["E0"], K1=pargs["K1"], K2=pargs["K2"], mu=pargs["mu"],
        kT=pargs["kT"], Fz=pargs["Fz"], Fx=pargs["Fx"], b=pargs["mlen"],
        chain_type=pargs["chain-type"], energy_type=pargs["energy-type"],
        phi_step=pargs["phi-step"], theta_step=pargs["theta-step"],
        adj_lb=pargs["step-adjust-lb"], adj_ub=pargs["step-adjust-ub"], adj_scale=pargs["step-adjust-scale"],
        steps_per_adjust=pargs["steps-per-adjust"], do_flips=pargs["do-flips"],
        umbrella=pargs["umbrella-sampling"], force_init=pargs["force-init"],
        accum_mode=0 if pargs.get("numeric-type", "float64") == "float64" else 1)


def validate(pargs: dict):
    """The reference's own refusals, same wording where it has one."""
    if pargs["acc"] != "metropolis":  # mcmc_eap_chain.jl:183-185
        raise lib.PolymcError(-1, f"'{pargs['acc']}' acceptance criteria has not yet been implemented.")
    if pargs["numeric-type"] not in ("float64", "float128", "dec128", "big"):  # :194-197
        raise lib.PolymcError(-1, f"numeric-type '{pargs['numeric-type']}' not understood")
    if pargs["profile"]:  # :378-380
        raise lib.PolymcError(-1, "not implemented for the HPC env")
    if pargs["ensemble-type"] != "force":
        # the reference warns that 'end-to-end' is experimental and unvalidated (:167-169); the
        # B200 path implements the fixed-force ensemble only and refuses instead of guessing.
        raise lib.PolymcError(-1, "ensemble-type '%s' is not supported by the B200 path (fixed-force only)"
                              % pargs["ensemble-type"])
    if pargs["num-steps"] < 0 or pargs["num-inits"] < 1 or pargs["replicas"] < 1:
        raise lib.PolymcError(-1, "num-steps, num-inits and replicas must be positive")


class Average:
    """Stand-in for StandardAverager / UmbrellaAverager (inc/average.jl:8-97) after the run:
    `get_avg` = value / normalizer (:38)."""

    def __init__(self, value, normalizer):
        self.value = value
        self.normalizer = normalizer

    def get_avg(self):
        return self.value / self.normalizer


def get_avg(a: Average):
    return a.get_avg()


def pool_replicas(sums: np.ndarray, umbrella: bool, extra: np.ndarray | None = None):
    """Pool the accumulators of R replica chains of one case: (values [17 (+2)], normaliser).

    Plain averagers (inc/average.jl:40-48): every replica has the same normaliser (its trial count), so
    Σ values / Σ normalisers is the mean over all trials of all replicas.  Umbrella averagers (:63-97): replica r's
    weights all carry the factor exp(log_gauge_r) with log_gauge_r = gauge0 + Ω0_r fixed at ITS initial chain
    (average.jl:109-118); Ω0 = Σ log sinθ has a spread of ~9 at n=100, so pooled sums would be dominated by one
    replica.  The gauge cancels inside each replica's own ratio value_r / normalizer_r (that ratio is the reference's
    estimate for one run), so the replicas are pooled as the mean of their ratios — what reduce_tabular_data.jl:36-56
    does with the `.out` files of repeated runs."""
    sums = np.asarray(sums, dtype=np.float64)
    if extra is not None:
        sums = np.concatenate([sums, np.asarray(extra, dtype=np.float64)], axis=1)
    if not umbrella:
        pooled = sums.sum(axis=0)
        return pooled, pooled[16]
    ratios = sums / sums[:, 16:17]
    pooled = ratios.mean(axis=0)          # pooled[16] == 1
    return pooled, 1.0


def mcmc(nsteps: int, pargs: dict):
    """`mcmc(nsteps, pargs)` of mcmc_eap_chain.jl:171-376.

    Returns (scalar_averagers, vector_averagers, ar) like the reference: scalar = [r², p², U, U²]
    (:243-249), vector = [r, r∘r, p, p∘p] (:250-255).  Writes <prefix>_trajectory.csv and
    <prefix>_rolling.csv (:256-259, :329-348) for replica 0.  With --replicas R > 1 the R chains run
    concurrently and their accumulators are pooled (Σ values / Σ normalisers)."""
    validate(pargs)
    seed = pargs.get("seed")
    if seed is None:  # the reference never seeds its RNG
        seed = time.time_ns() & 0xFFFFFFFFFFFF
    case = case_from_pargs(pargs)
    R = pargs["replicas"]
    stepout = pargs["stepout"]
    start = time.time()
    last_update = start
    with lib.Ensemble(case, replicas=R, seed=seed, device=pargs.get("device", 0)) as ens:
        ens.set_pair_precision(pargs.get("pair-precision", "fp64"))
        with open(f"{pargs['prefix']}_trajectory.csv", "w") as outfile, \
                open(f"{pargs['prefix']}_rolling.csv", "w") as rollfile:
            outfile.write(TRAJ_HEADER + "\n")
            rollfile.write(ROLL_HEADER + "\n")
            for init in range(1, pargs["num-inits"] + 1):
                # a few device launches per init keep the progress log (:294-299) alive on long runs
                chunk = nsteps
                if nsteps > 200000:
                    chunk = 200000 if stepout <= 0 else max(stepout, 200000 // stepout * stepout)
                done = 0
                while done < nsteps:
                    todo = min(chunk, nsteps - done)
                    traj, roll = ens.run(todo, stepout)
                    if traj is not None:
                        write_rows(outfile, traj[0])
                        write_rows(rollfile, roll[0])
                    done += todo
                    if time.time() - last_update > pargs["update-freq"]:
                        _log(pargs, "info", f"elapsed: {time.time() - start}")
                        _log(pargs, "info", f"init:    {init} / {pargs['num-inits']}")
                        _log(pargs, "info", f"step:    {done} / {nsteps}")
                        last_update = time.time()
                if init < pargs["num-inits"]:
                    ens.reinit()  # mcmc_eap_chain.jl:352-361 (after the last init it has no observable effect)
        sums = ens.accumulators()          # [R][17]
        diag = ens.diagnostics()
        if pargs.get("numeric-type", "float64") != "float64":
            # float128 | dec128 | big (:186-197): the device sums are double-double (value = hi + lo exactly); pool them
            # in exact rational arithmetic and keep the quotients for the printer
            from fractions import Fraction
            hi, lo = ens.accumulators_dd()
            per = [[Fraction(float(hi[r, k])) + Fraction(float(lo[r, k])) for k in range(17)] for r in range(R)]
            if pargs["umbrella-sampling"]:
                pargs["_extended"] = [sum(per[r][k] / per[r][16] for r in range(R)) / R for k in range(16)]
            else:
                tot = [sum(per[r][k] for r in range(R)) for k in range(17)]
                pargs["_extended"] = [tot[k] / tot[16] for k in range(16)]
    pooled, norm = pool_replicas(sums, pargs["umbrella-sampling"])
    ar = float(diag[:, 4].sum() / (R * pargs["num-inits"] * pargs["num-steps"])) if pargs["num-steps"] else 0.0
    _log(pargs, "info", f"total time elapsed: {time.time() - start}")
    _log(pargs, "info", f"acceptance rate: {ar}")
    vas = [Average(pooled[0:3].copy(), norm), Average(pooled[3:6].copy(), norm),
           Average(pooled[7:10].copy(), norm), Average(pooled[10:13].copy(), norm)]
    sas = [Average(pooled[6], norm), Average(pooled[13], norm), Average(pooled[14], norm), Average(pooled[15], norm)]
    return sas, vas, ar


def main(argv=None) -> int:
    """Top level of mcmc_eap_chain.jl (:155-165, :377-395): stdout carries ONLY the 10 result lines."""
    pargs = parse_args(argv)
    try:
        sas, vas, ar = mcmc(pargs["num-steps"], pargs)
    except lib.PolymcError as e:
        print(f"ERROR: {e}", file=sys.stderr)
        return 1
    avg16 = np.concatenate([vas[0].get_avg(), vas[1].get_avg(), [sas[0].get_avg()],
                            vas[2].get_avg(), vas[3].get_avg(), [sas[1].get_avg()],
                            [sas[2].get_avg(), sas[3].get_avg()]])
    if "_extended" in pargs:
        from .output import result_lines_extended
        lines = result_lines_extended(pargs["_extended"], ar, pargs["mlen"], pargs["num-monomers"], pargs["numeric-type"])
    else:
        lines = result_lines(avg16, ar, pargs["mlen"], pargs["num-monomers"])
    for line in lines:
        print(line)
    return 0
