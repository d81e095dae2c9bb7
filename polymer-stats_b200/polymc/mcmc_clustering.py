"""Host-side mirror of mcmc_clustering_eap_chain.jl: the ArgParse table (:19-153), the two
`mcmc(...)` methods (:166-352), the burn-in temperature ladder (:365-386) and the 12 result lines
(:389-400), driving libpolymc_b200.so through the C ABI (ABI v2 entry points pmc_begin_stage,
pmc_run_ex, pmc_init_x0, pmc_extra_averages).

Same option names, defaults, argument meaning and error behaviour as the reference.  Additive options
(not in the reference): --replicas, --seed, --device, --no-alpha-carry, --cutoff-full-energy.
"""
from __future__ import annotations

import argparse
import ast
import math
import operator
import sys
import time

import numpy as np

from . import lib
from .mcmc import Average, _log, pool_replicas
from .output import (ROLL_HEADER_CLUSTERING, julia_float, julia_vector, result_lines_clustering,
                     traj_header_clustering, write_rows)


def build_parser() -> argparse.ArgumentParser:
    """The @add_arg_table of mcmc_clustering_eap_chain.jl:19-153, option for option."""
    p = argparse.ArgumentParser(prog="mcmc_clustering_eap_chain", allow_abbrev=False,
                                description="fixed-force MCMC with cluster flips of an electro-active polymer "
                                            "chain (B200 path)")
    a = p.add_argument
    a("--E0", "-e", type=float, default=0.0, help="magnitude of the electric field")
    a("--chain-type", "-T", type=str, default="dielectric", help="chain type (dielectric|polar)")
    a("--K1", "-J", type=float, default=1.0, help="dipole susceptibility along the monomer axis")
    a("--K2", "-K", type=float, default=0.0, help="dipole susceptibility orthogonal to the monomer axis")
    a("--mu", "-m", type=float, default=1e-2, help="dipole magnitude (electret chain)")
    a("--bend-mod", "-a", type=float, default=0.0, help="bending modulus of chain")
    a("--bend-angle", "-g", type=float, default=0.0, help="zero energy bond angle")
    a("--energy-type", "-u", type=str, default="Ising",
      help="energy type (interacting|cutoff|Ising|noninteracting)")
    a("--cutoff-radius", type=float, default=7.5, help="cut off radius (units of monomer lengths)")
    a("--kT", "-k", type=float, default=1.0, help="dimensionless temperature")
    a("--Fz", "-F", type=float, default=0.0, help="force in the z-direction (direction of E-field)")
    a("--Fx", "-G", type=float, default=0.0, help="force in the x-direction")
    a("--mlen", "-b", type=float, default=1.0, help="monomer length")
    a("--num-monomers", "-n", type=int, default=100, help="number of monomers")
    a("--num-steps", "-N", type=int, default=1000000, help="number of steps")
    a("--phi-step", "-p", type=float, default=3 * math.pi / 8, help="maximum phi step length")
    a("--theta-step", "-q", type=float, default=3 * math.pi / 16, help="maximum theta step length")
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
    a("--burn-in", type=int, default=50000, help="steps for burn-in; i.e. steps before averaging")
    a("--burn-schedule", type=str, default="[1000; 100; 10; 2; 1]", help="temperature schedule for burn-in")
    a("--x0", type=str, default=None, help="initial configuration")
    a("--dx0", type=str, default="[2*pi, 1e-1]", help="random perturbation of x0")
    a("--profile", "-Z", action="store_true", help="profile the program")
    # additive
    a("--replicas", type=int, default=1, help="[B200 path] independent replica chains run concurrently and pooled")
    a("--seed", type=int, default=None, help="[B200 path] Philox seed (default: time-based, like the unseeded reference)")
    a("--device", type=int, default=0, help="[B200 path] CUDA device index")
    a("--no-alpha-carry", action="store_true",
      help="[B200 path] plain Metropolis-Hastings: do not keep log(alpha) in the acceptor's stored log-density "
           "(the reference keeps it, inc/acceptance.jl:30-33)")
    a("--cutoff-full-energy", action="store_true",
      help="[B200 path] with --energy-type cutoff, add the self energy and -r.F to the cut-off pair sum "
           "(the reference's UCutoff functor is the bare pair sum, inc/eap_chain.jl:171-192)")
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


# ---- `eval(Meta.parse(...))` of the vector-valued options (:366, inc/eap_chain.jl:64-65) -----------------
_BINOPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
           ast.Pow: operator.pow}


def _eval_num(node):
    if isinstance(node, ast.Expression):
        return _eval_num(node.body)
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return float(node.value)
    if isinstance(node, ast.Name) and node.id in ("pi", "π"):
        return math.pi
    if isinstance(node, ast.BinOp) and type(node.op) in _BINOPS:
        return _BINOPS[type(node.op)](_eval_num(node.left), _eval_num(node.right))
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
        v = _eval_num(node.operand)
        return -v if isinstance(node.op, ast.USub) else v
    raise ValueError("unsupported expression")


def parse_julia_vector(text: str, what: str) -> list:
    """A Julia vector literal of arithmetic expressions — `[1000; 100; 10; 2; 1]`, `[0.0; pi/2]`,
    `[2*pi, 1e-1]`, `[]` — as the reference `eval(Meta.parse(...))`s it.  Only numbers, pi/π and
    + - * / ^ are understood; anything else is refused (the reference would evaluate arbitrary Julia)."""
    t = text.strip()
    if not (t.startswith("[") and t.endswith("]")):
        raise lib.PolymcError(-1, f"Invalid input for '{what}', {text}")
    body = t[1:-1].strip()
    if not body:
        return []
    out = []
    for tok in body.replace(";", ",").split(","):
        tok = tok.strip().replace("^", "**")
        if not tok:
            continue
        try:
            out.append(_eval_num(ast.parse(tok, mode="eval")))
        except (ValueError, SyntaxError):
            raise lib.PolymcError(-1, f"Invalid input for '{what}', {text}") from None
    return out


def case_from_pargs(pargs: dict) -> lib.PmcCase:
    """EAPChain(pargs) argument mapping (inc/eap_chain.jl:60-135) for the clustering driver."""
    return lib.make_case(
        n=pargs["num-monomers"], E0=pargs["E0"], K1=pargs["K1"], K2=pargs["K2"], mu=pargs["mu"],
        kT=pargs["kT"], Fz=pargs["Fz"], Fx=pargs["Fx"], b=pargs["mlen"],
        chain_type=pargs["chain-type"], energy_type=pargs["energy-type"],
        phi_step=pargs["phi-step"], theta_step=pargs["theta-step"],
        adj_lb=pargs["step-adjust-lb"], adj_ub=pargs["step-adjust-ub"], adj_scale=pargs["step-adjust-scale"],
        steps_per_adjust=pargs["steps-per-adjust"], do_flips=False,
        umbrella=pargs["umbrella-sampling"], force_init=False,
        accum_mode=0 if pargs.get("numeric-type", "float64") == "float64" else 1,
        kappa=pargs["bend-mod"], psi0=pargs["bend-angle"], cutoff_radius=pargs["cutoff-radius"],
        cluster_prob=pargs["cluster-prob"], clustering=True,
        alpha_carry=not pargs.get("no-alpha-carry", False), cutoff_full=pargs.get("cutoff-full-energy", False))


def validate(pargs: dict):
    """The reference's own refusals, same wording where it has one."""
    if pargs["numeric-type"] not in ("float64", "float128", "dec128", "big"):  # :190-193
        raise lib.PolymcError(-1, f"numeric-type '{pargs['numeric-type']}' not understood")
    if pargs["profile"]:  # :354-358
        raise lib.PolymcError(-1, "Not currently implemented...")
    if pargs["num-steps"] < 0 or pargs["burn-in"] < 0 or pargs["replicas"] < 1:
        raise lib.PolymcError(-1, "num-steps, burn-in and replicas must be non-negative / positive")
    if pargs["num-monomers"] < 2:
        raise lib.PolymcError(-1, "num-monomers must be >= 2 (the <ψ> averager divides by n-1, :244)")


def dipoles_of(pargs: dict, phi: np.ndarray, theta: np.ndarray) -> np.ndarray:
    """chain.μs as a function of the angles (inc/dipole_response.jl:7-29): [..., n, 3]."""
    nh = np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)], axis=-1)
    if pargs["chain-type"] == "dielectric":
        mu = ((pargs["K1"] - pargs["K2"]) * pargs["E0"] * np.cos(theta))[..., None] * nh
        mu[..., 2] += pargs["K2"] * pargs["E0"]
        return mu
    return pargs["mu"] * nh


def _stage(ens, nsteps, pargs, kT_scale, write_files, start):
    """One `mcmc(nsteps, pargs, chain)` call (:171-352) on the chains held by `ens`."""
    ens.begin_stage(kT_scale)
    stepout = pargs["stepout"]
    n = pargs["num-monomers"]
    last_update = time.time()
    outfile = rollfile = None
    if write_files:  # every stage re-opens the files with "w" (:253,:258); only the last one survives
        outfile = open(f"{pargs['prefix']}_trajectory.csv", "w")
        rollfile = open(f"{pargs['prefix']}_rolling.csv", "w")
        outfile.write(traj_header_clustering(n) + "\n")
        rollfile.write(ROLL_HEADER_CLUSTERING + "\n")
    try:
        chunk = nsteps
        if nsteps > 200000:
            chunk = 200000 if stepout <= 0 else max(stepout, 200000 // stepout * stepout)
        done = 0
        while done < nsteps:
            todo = min(chunk, nsteps - done)
            traj, roll, state = ens.run_ex(todo, stepout if write_files else 0, fetch_rows=write_files,
                                           want_state=write_files)
            if write_files and traj is not None:
                st = state[0]                                     # [rows][2n] phi1,theta1,...
                mus = dipoles_of(pargs, st[:, 0::2], st[:, 1::2])  # [rows][n][3] → mux1,muy1,muz1,... (:318)
                write_rows(outfile, np.concatenate([traj[0], st, mus.reshape(len(st), -1)], axis=1))
                write_rows(rollfile, roll[0])
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
    """Top level of mcmc_clustering_eap_chain.jl:365-386: burn-in stages at kT × burn-schedule, then the
    production stage.  Returns (scalar_averagers[6], vector_averagers[4], ar) of the production stage."""
    validate(pargs)
    kT_multipliers = parse_julia_vector(pargs["burn-schedule"], "burn-schedule")
    seed = pargs.get("seed")
    if seed is None:  # the reference never seeds its RNG
        seed = time.time_ns() & 0xFFFFFFFFFFFF
    case = case_from_pargs(pargs)
    R = pargs["replicas"]
    start = time.time()
    with lib.Ensemble(case, replicas=R, seed=seed, device=pargs.get("device", 0)) as ens:
        if pargs.get("x0") is not None:  # inc/eap_chain.jl:63-78
            x0 = parse_julia_vector(pargs["x0"], "x0")
            dx0 = parse_julia_vector(pargs["dx0"], "dx0")
            if len(x0) not in (2, 2 * pargs["num-monomers"]) or len(dx0) < 2:
                raise lib.PolymcError(-1, f"Invalid input for 'x0', {pargs['x0']}")
            ens.init_x0(x0, dx0[:2])
        for mult in kT_multipliers:  # :367-381
            _stage(ens, pargs["burn-in"], pargs, mult, write_files=False, start=start)
        _stage(ens, pargs["num-steps"], pargs, 1.0, write_files=True, start=start)  # :383-384
        sums = ens.accumulators()        # [R][17]
        xsums = ens.extra_accumulators()  # [R][2]
        diag = ens.diagnostics()
    pooled, norm = pool_replicas(sums, pargs["umbrella-sampling"], extra=xsums)
    xpooled = pooled[17:19]
    ar = float(diag[:, 4].sum() / (R * pargs["num-steps"])) if pargs["num-steps"] else 0.0  # :340
    _log(pargs, "info", f"total time elapsed: {time.time() - start}")
    _log(pargs, "info", f"acceptance rate: {ar}")
    vas = [Average(pooled[0:3].copy(), norm), Average(pooled[3:6].copy(), norm),
           Average(pooled[7:10].copy(), norm), Average(pooled[10:13].copy(), norm)]
    sas = [Average(pooled[6], norm), Average(pooled[13], norm), Average(pooled[14], norm), Average(pooled[15], norm),
           Average(xpooled[0], norm), Average(xpooled[1], norm)]
    return sas, vas, ar


def main(argv=None) -> int:
    """stdout carries ONLY the 12 result lines (:389-400)."""
    pargs = parse_args(argv)
    try:
        sas, vas, ar = mcmc_ladder(pargs)
    except lib.PolymcError as e:
        print(f"ERROR: {e}", file=sys.stderr)
        return 1
    avg16 = np.concatenate([vas[0].get_avg(), vas[1].get_avg(), [sas[0].get_avg()],
                            vas[2].get_avg(), vas[3].get_avg(), [sas[1].get_avg()],
                            [sas[2].get_avg(), sas[3].get_avg()]])
    for line in result_lines_clustering(avg16, sas[4].get_avg(), sas[5].get_avg(), ar, pargs["mlen"],
                                        pargs["num-monomers"]):
        print(line)
    return 0


__all__ = ["build_parser", "parse_args", "default_pargs", "parse_julia_vector", "case_from_pargs", "validate",
           "dipoles_of", "mcmc_ladder", "main", "julia_float", "julia_vector"]
