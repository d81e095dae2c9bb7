"""Host-side mirror of the 2-D tree's driver, 2D/mcmc_clustering_eap_chain.jl: the
ArgParse table (:19-133), `mcmc(...)` (:143-311), the burn-in ladder (:323-336) and the 10 result lines with
2-vectors (:338-347), on libpolymc_b200.so with `planar = 1` cases.

The planar chain has the state ϕ only: n̂ = (cosϕ, sinϕ) with the field along the second axis
(2D/inc/eap_chain.jl:33, 2D/inc/dipole_response.jl:7-10); it maps onto the 3-D kernels as the x–z plane
(n̂y = 0, sinθ ≡ 1).  Reference behaviour kept: every `mcmc()` call — each burn-in stage and the production run
— starts from a NEW random chain (`chain = EAPChain(pargs)`, :151), so the ladder does not carry the chain.
"""
from __future__ import annotations

import argparse
import math
import sys
import time

from . import lib
from .mcmc import Average, _log, pool_replicas
from .mcmc_clustering import parse_julia_vector
from .output import julia_float, julia_vector, write_rows

TRAJ_HEADER_2D = "step,r1,r3,p1,p3,U"                                                  # :228
ROLL_HEADER_2D = "step,r1,r3,r1sq,r3sq,rsq,p1,p3,p1sq,p3sq,psq,U,Usq"                  # :230
# columns of the 3-D rows that make the 2-D rows (x ↔ "1", z ↔ "3")
_TRAJ_COLS = [0, 1, 3, 4, 6, 7]
_ROLL_COLS = [0, 1, 3, 4, 6, 7, 8, 10, 11, 13, 14, 15, 16]


def build_parser() -> argparse.ArgumentParser:
    """The @add_arg_table of 2D/mcmc_clustering_eap_chain.jl:19-133, option for option."""
    p = argparse.ArgumentParser(prog="mcmc_clustering_eap_chain_2d", allow_abbrev=False,
                                description="fixed-force MCMC with cluster flips of a planar electro-active polymer "
                                            "chain (B200 path)")
    a = p.add_argument
    a("--E0", "-e", type=float, default=0.0, help="magnitude of the electric field")
    a("--chain-type", "-T", type=str, default="dielectric", help="chain type (dielectric|polar)")
    a("--K1", "-J", type=float, default=1.0, help="dipole susceptibility along the monomer axis")
    a("--K2", "-K", type=float, default=0.0, help="dipole susceptibility orthogonal to the monomer axis")
    a("--mu", "-m", type=float, default=1e-2, help="dipole magnitude (electret chain)")
    a("--energy-type", "-u", type=str, default="noninteracting", help="energy type (noninteracting|interacting|Ising)")
    a("--kT", "-k", type=float, default=1.0, help="dimensionless temperature")
    a("--Fz", "-F", type=float, default=0.0, help="force in the z-direction (direction of E-field)")
    a("--Fx", "-G", type=float, default=0.0, help="force in the x-direction")
    a("--mlen", "-b", type=float, default=1.0, help="monomer length")
    a("--num-monomers", "-n", type=int, default=100, help="number of monomers")
    a("--num-steps", "-N", type=int, default=1000000, help="number of steps")
    a("--phi-step", "-p", type=float, default=3 * math.pi / 8, help="maximum phi step length")
    a("--cluster-prob", type=float, default=0.5, help="probability of flipping a cluster")
    a("--step-adjust-lb", "-L", type=float, default=0.15, help="lower acceptance bound for step adaptation")
    a("--step-adjust-ub", "-U", type=float, default=0.40, help="upper acceptance bound for step adaptation")
    a("--step-adjust-scale", "-A", type=float, default=1.1, help="step adaptation factor (1.0 disables)")
    a("--steps-per-adjust", "-S", type=int, default=2500, help="steps between step-size adjustments")
    a("--umbrella-sampling", "-B", action="store_true", help="use umbrella sampling")
    a("--update-freq", type=float, default=15.0, help="progress update frequency (seconds)")
    a("--verbose", "-v", type=int, default=3, help="verbosity 0-3")
    a("--prefix", "-P", type=str, default="eap-mcmc", help="prefix for output files")
    a("--postfix", "-Q", type=str, default="", help="postfix for output files (parsed, unused)")
    a("--stepout", "-s", type=int, default=500, help="steps between storing microstates")
    a("--numeric-type", type=str, default="float64", help="accumulator type (float64|float128|dec128|big)")
    a("--profile", "-Z", action="store_true", help="profile the program")
    a("--burn-in", type=int, default=50000, help="steps for burn-in; i.e. steps before averaging")
    a("--burn-schedule", type=str, default="[1000; 100; 10; 2; 1]", help="temperature schedule for burn-in")
    # additive
    a("--replicas", type=int, default=1, help="[B200 path] independent replica chains run concurrently and pooled")
    a("--seed", type=int, default=None, help="[B200 path] Philox seed (default: time-based, like the unseeded reference)")
    a("--device", type=int, default=0, help="[B200 path] CUDA device index")
    a("--no-alpha-carry", action="store_true",
      help="[B200 path] plain Metropolis-Hastings (the reference keeps log(alpha) in the acceptor, "
           "2D/inc/acceptance.jl:30-33)")
    return p


def parse_args(argv=None) -> dict:
    ns = build_parser().parse_args(argv)
    return {k.replace("_", "-"): v for k, v in vars(ns).items()}


def default_pargs(**overrides) -> dict:
    d = parse_args([])
    for k, v in overrides.items():
        d[k.replace("_", "-")] = v
    return d


def case_from_pargs(pargs: dict) -> lib.PmcCase:
    """EAPChain(pargs) argument mapping of the 2-D tree (2D/inc/eap_chain.jl:66-112)."""
    if pargs["energy-type"] == "cutoff":
        raise lib.PolymcError(-1, "energy-type is not understood.")   # 2D/inc/eap_chain.jl:88-90
    return lib.make_case(
        n=pargs["num-monomers"], E0=pargs["E0"], K1=pargs["K1"], K2=pargs["K2"], mu=pargs["mu"],
        kT=pargs["kT"], Fz=pargs["Fz"], Fx=pargs["Fx"], b=pargs["mlen"],
        chain_type=pargs["chain-type"], energy_type=pargs["energy-type"],
        # there is no θ: θstep is tied to ϕstep so that the shared adaptation rule caps at ϕstep = π exactly
        # as `ϕstep != π` does (2D/mcmc_clustering_eap_chain.jl:262-273)
        phi_step=pargs["phi-step"], theta_step=pargs["phi-step"] / 2,
        adj_lb=pargs["step-adjust-lb"], adj_ub=pargs["step-adjust-ub"], adj_scale=pargs["step-adjust-scale"],
        steps_per_adjust=pargs["steps-per-adjust"], umbrella=pargs["umbrella-sampling"],
        accum_mode=0 if pargs.get("numeric-type", "float64") == "float64" else 1,
        cluster_prob=pargs["cluster-prob"], clustering=True, alpha_carry=not pargs.get("no-alpha-carry", False),
        planar=True)


def validate(pargs: dict):
    if pargs["numeric-type"] not in ("float64", "float128", "dec128", "big"):
        raise lib.PolymcError(-1, f"numeric-type '{pargs['numeric-type']}' not understood")
    if pargs["profile"]:
        raise lib.PolymcError(-1, "Not currently implemented...")
    if pargs["num-steps"] < 0 or pargs["burn-in"] < 0 or pargs["replicas"] < 1 or pargs["num-monomers"] < 2:
        raise lib.PolymcError(-1, "num-steps, burn-in must be >= 0, replicas >= 1, num-monomers >= 2")


def _stage(ens, nsteps, pargs, kT_scale, write_files, start):
    ens.begin_stage(kT_scale)   # planar handles draw a NEW random chain here (2D/...:151)
    stepout = pargs["stepout"] if write_files else 0
    outfile = rollfile = None
    if write_files:
        outfile = open(f"{pargs['prefix']}_trajectory.csv", "w")
        rollfile = open(f"{pargs['prefix']}_rolling.csv", "w")
        outfile.write(TRAJ_HEADER_2D + "\n")
        rollfile.write(ROLL_HEADER_2D + "\n")
    last_update = time.time()
    try:
        chunk = nsteps
        if nsteps > 200000:
            chunk = 200000 if stepout <= 0 else max(stepout, 200000 // stepout * stepout)
        done = 0
        while done < nsteps:
            todo = min(chunk, nsteps - done)
            traj, roll = ens.run(todo, stepout, fetch_rows=write_files)
            if write_files and traj is not None:
                write_rows(outfile, traj[0][:, _TRAJ_COLS])
                write_rows(rollfile, roll[0][:, _ROLL_COLS])
            done += todo
            if time.time() - last_update > pargs["update-freq"]:
                _log(pargs, "info", f"elapsed: {time.time() - start}")
                _log(pargs, "info", f"step:    {done} / {nsteps}")
                last_update = time.time()
    finally:
        if outfile:
            outfile.close()
            rollfile.close()


def mcmc_ladder(pargs: dict):
    """Top level of 2D/mcmc_clustering_eap_chain.jl:323-336.  Returns (scalar_averagers[4], vector_averagers[4]
    of 2-vectors, ar) of the production stage."""
    validate(pargs)
    kT_multipliers = parse_julia_vector(pargs["burn-schedule"], "burn-schedule")
    if not kT_multipliers:
        raise lib.PolymcError(-1, "burn-schedule must not be empty (kT_multipliers[1], 2D/...:326)")
    seed = pargs.get("seed")
    if seed is None:
        seed = time.time_ns() & 0xFFFFFFFFFFFF
    case = case_from_pargs(pargs)
    R = pargs["replicas"]
    start = time.time()
    with lib.Ensemble(case, replicas=R, seed=seed, device=pargs.get("device", 0)) as ens:
        for mult in kT_multipliers:
            _stage(ens, pargs["burn-in"], pargs, mult, write_files=False, start=start)
        _stage(ens, pargs["num-steps"], pargs, 1.0, write_files=True, start=start)
        sums = ens.accumulators()
        diag = ens.diagnostics()
    pooled, norm = pool_replicas(sums, pargs["umbrella-sampling"])
    ar = float(diag[:, 4].sum() / (R * pargs["num-steps"])) if pargs["num-steps"] else 0.0
    _log(pargs, "info", f"total time elapsed: {time.time() - start}")
    _log(pargs, "info", f"acceptance rate: {ar}")
    xz = [0, 2]
    vas = [Average(pooled[0:3][xz].copy(), norm), Average(pooled[3:6][xz].copy(), norm),
           Average(pooled[7:10][xz].copy(), norm), Average(pooled[10:13][xz].copy(), norm)]
    sas = [Average(pooled[6], norm), Average(pooled[13], norm), Average(pooled[14], norm), Average(pooled[15], norm)]
    return sas, vas, ar


def result_lines_2d(sas, vas, ar, mlen, n):
    """The 10 stdout lines of the 2-D driver (:338-347): vectors have two components."""
    nb = mlen * n
    r = vas[0].get_avg()
    return [
        f"<r>    =   {julia_vector(r)}",
        f"<r/nb> =   {julia_vector([x / nb for x in r])}",
        f"<rj2>  =   {julia_vector(vas[1].get_avg())}",
        f"<r2>   =   {julia_float(sas[0].get_avg())}",
        f"<p>    =   {julia_vector(vas[2].get_avg())}",
        f"<pj2>  =   {julia_vector(vas[3].get_avg())}",
        f"<p2>   =   {julia_float(sas[1].get_avg())}",
        f"<U>    =   {julia_float(sas[2].get_avg())}",
        f"<U2>   =   {julia_float(sas[3].get_avg())}",
        f"AR     =   {julia_float(ar)}",
    ]


def main(argv=None) -> int:
    pargs = parse_args(argv)
    try:
        sas, vas, ar = mcmc_ladder(pargs)
    except lib.PolymcError as e:
        print(f"ERROR: {e}", file=sys.stderr)
        return 1
    for line in result_lines_2d(sas, vas, ar, pargs["mlen"], pargs["num-monomers"]):
        print(line)
    return 0
