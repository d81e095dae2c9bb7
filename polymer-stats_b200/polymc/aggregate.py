"""Batched aggregation: the tables that scripts/aggregate_mcmc.jl and
scripts/reduce_tabular_data.jl build from thousands of per-case `<prefix>.out` files, emitted directly
from the gathered device averages of a sweep.

The reference's pipeline is: launcher (run/*.jl) names each case `K-V_K-V_…[_run-NNN]` with
V = @sprintf("%07d", round(Int, 1e3*x)) (run/interacting_dielectric_study.jl:12-17), captures the stdout of
one `julia mcmc_*.jl` process into `<prefix>.out` (:41-43); aggregate_mcmc.jl then turns every file NAME
into the parameter columns (`eval(Meta.parse(V))*1e-3`, :63-70) and every LINE into output columns in line
order (:71-72), one CSV row per file in `readdir` (= sorted) order (:61); reduce_tabular_data.jl pools rows
with identical parameter columns (runs) by averaging (:36-56).  With one GPU process running the whole
sweep, formatting thousands of `.out` files and re-parsing them is the wall-clock bottleneck, so the same
rows are produced in memory here; writing the per-case `.out` files stays available (`write_out_files`).
"""
from __future__ import annotations

import os

import numpy as np

from .output import julia_float, result_lines, result_lines_2d, result_lines_clustering

# aggregate_mcmc.jl:40-47 (+ "kappa" :50-52)
INPUT_HEADERS = {"dielectric": ["E0", "K1", "K2", "kT", "Fz", "Fx", "n", "b"],
                 "polar": ["E0", "mu", "kT", "Fz", "Fx", "n", "b"]}
# aggregate_mcmc.jl:54-58
OUTPUT_HEADERS_3D = ["r1", "r2", "r3", "lambda1", "lambda2", "lambda3", "r1sq", "r2sq", "r3sq", "rsquared",
                     "p1", "p2", "p3", "p1sq", "p2sq", "p3sq", "psquared", "U", "Usquared", "Ealign", "psi", "AR"]
# aggregate_mcmc.jl:56-57 (dims == 2): the 10 lines of the 2-D driver, vectors with two components
OUTPUT_HEADERS_2D = ["r1", "r2", "lambda1", "lambda2", "r1sq", "r2sq", "rsquared", "p1", "p2", "p1sq", "p2sq",
                     "psquared", "U", "Usquared", "AR"]
# pargs key of each file-name token
_PARG_OF = {"E0": "E0", "K1": "K1", "K2": "K2", "mu": "mu", "kT": "kT", "Fz": "Fz", "Fx": "Fx", "n": "num-monomers",
            "b": "mlen", "kappa": "bend-mod"}


def fmt(x) -> str:
    """`fmt(x) = @sprintf("%07d", round(Int, 1e3*x))` (run/*.jl; Julia rounds half away from zero)."""
    v = 1e3 * float(x)
    r = int(np.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)
    return f"{r:07d}"


def fmt_int(x) -> str:
    return f"{int(x):03d}"


def input_headers(chain_type: str, kappaflag: bool = False):
    if chain_type not in INPUT_HEADERS:
        raise ValueError("I don't understand the chain type")     # aggregate_mcmc.jl:45-46
    return INPUT_HEADERS[chain_type] + (["kappa"] if kappaflag else [])


def prefix_of(pargs: dict, chain_type: str, kappaflag: bool = False, run=None) -> str:
    """The launchers' `prefix(case)` (e.g. run/phases-kT-small-n_2023-09-09.jl:14-16,
    run/Ising_2025-12-17.jl:14-16): tokens in the order aggregate_mcmc.jl's headers expect."""
    toks = [f"{k}-{fmt(pargs[_PARG_OF[k]])}" for k in input_headers(chain_type, kappaflag)]
    if run is not None:
        toks.append(f"run-{fmt_int(run)}")
    return "_".join(toks)


def params_from_prefix(prefix: str, runflag: bool = False):
    """aggregate_mcmc.jl:62-70: `split(basename, ".")[1]`, split on "_", optional `pop!` of the run token,
    then `eval(Meta.parse(V))*1e-3` of everything after the first "-"."""
    fields = os.path.basename(prefix).split(".")[0].split("_")
    if runflag:
        fields.pop()
    return [int(f.split("-", 1)[1]) * 1e-3 for f in fields]


def output_values(avg16, acc_rate, mlen, n, extras=None):
    """The output columns of one case in stdout line order (aggregate_mcmc.jl:71-72): r, r/nb, rj2, r2, p, pj2,
    p2, U, U2, [cos2, psi,] AR — 20 values for mcmc_eap_chain, 22 for the clustering driver."""
    nb = mlen * n
    v = list(avg16[0:3]) + [x / nb for x in avg16[0:3]] + list(avg16[3:6]) + [avg16[6]]
    v += list(avg16[7:10]) + list(avg16[10:13]) + [avg16[13], avg16[14], avg16[15]]
    if extras is not None:
        v += [extras[0], extras[1]]
    return [float(x) for x in v] + [float(acc_rate)]


def output_values_2d(avg16, acc_rate, mlen, n):
    """The 15 output columns of one 2-D case in stdout line order (2D/mcmc_clustering_eap_chain.jl:338-347)."""
    nb = mlen * n
    r, rj2, p, pj2 = ([a[0], a[2]] for a in (avg16[0:3], avg16[3:6], avg16[7:10], avg16[10:13]))
    v = r + [x / nb for x in r] + rj2 + [avg16[6]] + p + pj2 + [avg16[13], avg16[14], avg16[15]]
    return [float(x) for x in v] + [float(acc_rate)]


def out_text_2d(avg16, acc_rate, mlen, n) -> str:
    return "\n".join(result_lines_2d(avg16, acc_rate, mlen, n)) + "\n"


def out_text(avg16, acc_rate, mlen, n, extras=None) -> str:
    """Content of `<prefix>.out`: exactly the stdout of the driver."""
    lines = (result_lines(avg16, acc_rate, mlen, n) if extras is None
             else result_lines_clustering(avg16, extras[0], extras[1], acc_rate, mlen, n))
    return "\n".join(lines) + "\n"


def aggregate_table(entries, chain_type: str, kappaflag: bool = False, runflag: bool = False, dims: int = 3):
    """entries: iterable of (prefix, output value list).  Returns (header list, rows) exactly as
    aggregate_mcmc.jl writes them: one row per file in sorted file-name order, parameter columns parsed back
    from the NAME (so they carry the launchers' 1e-3 rounding), then the output columns (dims = 2: the 2-D
    tree's, aggregate_mcmc.jl:23-35,56-57)."""
    header = input_headers(chain_type, kappaflag) + (OUTPUT_HEADERS_3D if dims == 3 else OUTPUT_HEADERS_2D)
    rows = []
    for prefix, values in sorted(entries, key=lambda e: os.path.basename(e[0]) + ".out"):
        rows.append(params_from_prefix(prefix, runflag) + list(values))
    return header, rows


def _by_pattern(name: str, param: str, runflag: bool):
    """The glob pattern scripts/aggregate_by.jl builds from one file NAME (:40-57), or None when the name does not
    carry `param` (:41-42).  param "FxFz" sweeps both force components (:14-18,48-51)."""
    if param == "FxFz":
        start = name.find("Fz-")                                                # :49
        fx = name.find("Fx-")
        if start < 0 or fx < 0:
            return None
        us = name.find("_", fx + 3)                                             # :50 (the "_" after the Fx value)
        pattern = name[:start] + "Fz-*_Fx-*" + (name[us:] if us >= 0 else "")
    else:
        at = name.find(param + "-")                                             # :40
        if at < 0:
            return None
        vstart = at + len(param) + 1                                            # :43
        us = name.find("_", vstart)                                             # :44
        vend = us if us >= 0 else len(name) - 4                                 # :45 (`length(datafile)-3`, 1-based)
        pattern = name[:vstart] + "*" + name[vend:]                             # :52
    if runflag:
        pattern = pattern[:-5] + "*" + pattern[-4:]                             # :54 (see aggregate_by)
    return pattern


def aggregate_by(entries, param: str, chain_type: str, kappaflag: bool = False, runflag: bool = False,
                 dims: int = 3):
    """scripts/aggregate_by.jl:25-60 in memory: one table per combination of the parameters other than `param`
    — the sweep over `param` at fixed everything else — as {"<other tokens joined by _>.csv": (header, rows)}.

    Walks the file names in sorted order; the first name of each group defines a glob pattern with `param`'s
    value (and, with runflag, the LAST DIGIT of the run number, :54) wildcarded, and every name matching it goes
    through aggregate_mcmc.jl.  Upstream behaviour kept: with runflag the pattern is `…_run-00*.out`, so only
    runs 000-009 of a group reach its table."""
    import fnmatch
    entries = sorted(entries, key=lambda e: os.path.basename(e[0]) + ".out")
    names = [os.path.basename(e[0]) + ".out" for e in entries]
    tables, seen = {}, []
    for name in names:
        fileparams = "".join(name.split(".")[:-1]).split("_")                   # :28
        if runflag:
            fileparams.pop()                                                    # :29
        if param == "FxFz":
            filtered = [t for t in fileparams if not t.startswith("Fz") and not t.startswith("Fx")]   # :31
        else:
            filtered = [t for t in fileparams if not t.startswith(param)]       # :33
        if filtered in seen:                                                    # :35-37
            continue
        seen.append(filtered)
        pattern = _by_pattern(name, param, runflag)
        if pattern is None:
            continue
        chosen = [e for e, nm in zip(entries, names) if fnmatch.fnmatchcase(nm, pattern)]
        tables["_".join(filtered) + ".csv"] = aggregate_table(chosen, chain_type, kappaflag, runflag, dims)   # :56-58
    return tables


def reduce_table(header, rows, nparams: int):
    """reduce_tabular_data.jl:36-56: pool rows with identical parameter columns — plain mean over the runs of
    every output column — sorted by the parameter tuple."""
    pooled = {}
    for row in rows:
        k = tuple(row[:nparams])
        acc = pooled.setdefault(k, [np.zeros(len(row) - nparams), 0])
        acc[0] = acc[0] + np.asarray(row[nparams:], dtype=float)
        acc[1] += 1
    out = [list(k) + list(pooled[k][0] / pooled[k][1]) for k in sorted(pooled)]
    return header, out


def write_table(path: str, header, rows):
    """`writedlm(outfile, …, ',')` (aggregate_mcmc.jl:59,66-75): Float64 printed the way Julia prints them."""
    with open(path, "w") as f:
        f.write(",".join(header) + "\n")
        for row in rows:
            f.write(",".join(julia_float(x) for x in row) + "\n")


def write_out_files(outdir: str, entries_text):
    """Optional per-case `<prefix>.out` files (what the launchers write, run/*.jl `write(outfile, output)`)."""
    os.makedirs(outdir, exist_ok=True)
    for prefix, text in entries_text:
        with open(os.path.join(outdir, os.path.basename(prefix) + ".out"), "w") as f:
            f.write(text)
