"""Restatement of scripts/aggregate_mcmc.jl, scripts/aggregate_by.jl and scripts/reduce_tabular_data.jl working on `.out` FILES, the
way the reference does (TEST INFRASTRUCTURE ONLY — the checker for polymc.aggregate, which builds the same
tables in memory).  Each step cites the script line it follows."""
from __future__ import annotations

import ast
import fnmatch
import os

INPUT_HEADERS = {"dielectric": ["E0", "K1", "K2", "kT", "Fz", "Fx", "n", "b"],     # aggregate_mcmc.jl:40-41
                 "polar": ["E0", "mu", "kT", "Fz", "Fx", "n", "b"]}               # :42-43
OUTPUT_HEADERS_3D = ["r1", "r2", "r3", "lambda1", "lambda2", "lambda3", "r1sq", "r2sq", "r3sq", "rsquared",
                     "p1", "p2", "p3", "p1sq", "p2sq", "p3sq", "psquared", "U", "Usquared", "Ealign", "psi", "AR"]  # :54-55


def _julia_eval(text: str):
    """eval(Meta.parse(...)) of a result value: a Float64 literal or a `[a, b, c]` vector (:71)."""
    t = text.strip().replace("NaN", "float('nan')").replace("Inf", "float('inf')")
    v = eval(t, {"__builtins__": {}, "float": float})  # noqa: S307 - test infrastructure, our own files
    return list(v) if isinstance(v, (list, tuple)) else [v]


OUTPUT_HEADERS_2D = ["r1", "r2", "lambda1", "lambda2", "r1sq", "r2sq", "rsquared", "p1", "p2", "p1sq", "p2sq",
                     "psquared", "U", "Usquared", "AR"]                                                           # :56-57


def aggregate_mcmc(indir: str, pattern: str, chain_type: str, kappaflag=False, runflag=False, dims=3):
    """aggregate_mcmc.jl:36-77 → (header, rows)."""
    header = list(INPUT_HEADERS[chain_type])
    if kappaflag:
        header.append("kappa")                                                  # :50-52
    header += OUTPUT_HEADERS_3D if dims == 3 else OUTPUT_HEADERS_2D             # :53-57
    rows = []
    for name in sorted(os.listdir(indir)):                                      # readdir(pattern, indir), :61
        if not fnmatch.fnmatchcase(name, pattern):
            continue
        fields = name.split(".")[0].split("_")                                   # :63
        if runflag:
            fields.pop()                                                        # :64-66
        params = [ast.literal_eval(f.split("-", 1)[1].lstrip("0") or "0") * 1e-3 for f in fields]   # :69-70
        vals = []
        with open(os.path.join(indir, name), encoding="utf-8") as fh:
            for line in fh.read().splitlines():                                  # readlines(infile), :72
                vals += _julia_eval(line.split("=")[1])                          # :71
        rows.append(params + vals)
    return header, rows


def reduce_tabular_data(header, rows, chain_type: str, kappaflag=False):
    """reduce_tabular_data.jl:31-59 on one table."""
    nparams = len(INPUT_HEADERS[chain_type]) + (1 if kappaflag else 0)          # :16-27
    pooled = {}
    for row in rows:                                                            # :38-50
        k = tuple(row[:nparams])
        data = [x for x in row[nparams:] if x != ""]
        if k in pooled:
            pooled[k] = ([a + b for a, b in zip(pooled[k][0], data)], pooled[k][1] + 1)
        else:
            pooled[k] = (data, 1)
    out = []
    for k in sorted(pooled):                                                    # :52-56
        v, cnt = pooled[k]
        out.append(list(k) + [x / cnt for x in v])
    return header, out


def aggregate_by(indir: str, param_arg: str, chain_type: str, kappaflag=False, runflag=False, dims=3):
    """aggregate_by.jl:11-60 → {basename of the CSV it would write: (header, rows)}; Julia's 1-based string
    indices are kept literally (helper `J`) so that the slicing quirks carry over."""
    fxfz = param_arg == "FxFz"                                                   # :14-18
    param = "Fz" if fxfz else param_arg

    def find(hay, needle, start1):   # findnext(needle, hay, start) → (first, last) 1-based, or None
        k = hay.find(needle, start1 - 1)
        return None if k < 0 else (k + 1, k + len(needle))

    def J(text, a, b):               # text[a:b], 1-based inclusive
        return text[a - 1:b] if b >= a else ""

    out, params_ran = {}, []
    datafiles = sorted(f for f in os.listdir(indir) if fnmatch.fnmatchcase(f, "*.out"))   # :24
    for datafile in datafiles:
        fileparams = "".join(datafile.split(".")[:-1]).split("_")                # :28
        if runflag:
            fileparams.pop()                                                    # :29
        if fxfz:
            filtered = [x for x in fileparams if not x.startswith("Fz") and not x.startswith("Fx")]   # :31
        else:
            filtered = [x for x in fileparams if not x.startswith(param)]       # :33
        if filtered in params_ran:                                              # :35-39
            continue
        params_ran.append(filtered)
        strspan = find(datafile, param + "-", 1)                                # :41
        if strspan is None:
            continue                                                            # :42
        value_start = strspan[1] + 1                                            # :43
        strspan2 = find(datafile, "_", value_start)                             # :44
        value_end = strspan2[0] if strspan2 is not None else len(datafile) - 3   # :45
        if fxfz:
            start_index = find(datafile, "Fz-", 1)[0]                           # :49
            end_index = find(datafile, "_", find(datafile, "Fx-", 1)[1])[1] - 1   # :50
            pattern = J(datafile, 1, start_index - 1) + "Fz-*_Fx-*" + J(datafile, end_index + 1, len(datafile))   # :51
        else:
            pattern = J(datafile, 1, value_start - 1) + "*" + J(datafile, value_end, len(datafile))   # :53
        if runflag:
            pattern = J(pattern, 1, len(pattern) - 5) + "*" + J(pattern, len(pattern) - 3, len(pattern))   # :55
        out["_".join(filtered) + ".csv"] = aggregate_mcmc(indir, pattern, chain_type, kappaflag, runflag, dims)  # :57-59
    return out
