"""The output contract of mcmc_eap_chain.jl: the 10 `key = value` stdout lines
(:386-395) and the two CSV files (:256-259, :329-348), formatted the way Julia prints Float64 so
run/*.jl launchers, scripts/aggregate_mcmc.jl (:61-75) and scripts/plot_hermans.py keep parsing them.
"""
from __future__ import annotations

import math
from decimal import Decimal

TRAJ_HEADER = "step,r1,r2,r3,p1,p2,p3,U"                                   # mcmc_eap_chain.jl:257
ROLL_HEADER = ("step,r1,r2,r3,r1sq,r2sq,r3sq,rsq,p1,p2,p3,p1sq,p2sq,p3sq,psq,U,Usq")  # :259


def julia_float(x: float) -> str:
    """Julia `print(::Float64)`: shortest round-trip digits; positional notation for decimal
    exponents with -4 < pt <= 6 (Base.Ryu.writeshortest), otherwise `d.ddde±x`; `NaN`, `Inf`, `-Inf`."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    digs = "".join(map(str, digits)).rstrip("0") or "0"
    exp += len(digits) - len(digs)          # value = 0.digs… × 10^(len(digs)+exp) → point position
    pt = len(digs) + exp                    # number of digits before the decimal point
    s = "-" if sign else ""
    if -4 < pt <= 6:
        if pt <= 0:
            body = "0." + "0" * (-pt) + digs
        elif pt >= len(digs):
            body = digs + "0" * (pt - len(digs)) + ".0"
        else:
            body = digs[:pt] + "." + digs[pt:]
        return s + body
    mant = digs[0] + "." + (digs[1:] or "0")
    return f"{s}{mant}e{pt - 1}"


EXT_DIGITS = {"float128": 36, "dec128": 34, "big": 77}   # significant digits of Quadmath / DecFP / BigFloat(256) text


def extended_text(x, numeric_type: str) -> str:
    """A `fractions.Fraction` (the exact quotient of double-double sums) as Float128 / Dec128 / BigFloat would print
    it: `d.ddd…e±XX` with the type's number of significant digits (mcmc_eap_chain.jl:186-197 selects the type)."""
    from decimal import Decimal, localcontext
    digits = EXT_DIGITS[numeric_type]
    with localcontext() as ctx:
        ctx.prec = digits + 10
        d = Decimal(x.numerator) / Decimal(x.denominator)
        if d == 0:
            return "0.0"
        ctx.prec = digits
        d = +d
    sign, dg, exp = d.as_tuple()
    dg = "".join(map(str, dg))
    e10 = exp + len(dg) - 1
    mant = dg[0] + "." + (dg[1:].rstrip("0") if numeric_type == "big" else dg[1:].ljust(digits - 1, "0")) 
    if mant.endswith("."):
        mant += "0"
    return f"{'-' if sign else ''}{mant}e{'+' if e10 >= 0 else '-'}{abs(e10):02d}"


def result_lines_extended(fr16, acc_rate, mlen, n, numeric_type):
    """The 10 stdout lines with --numeric-type float128|dec128|big: the averages are exact quotients of the device's
    double-double sums (pmc_accumulators_dd), printed with the digits of the chosen type."""
    from fractions import Fraction
    t = lambda x: extended_text(x, numeric_type)
    vec = lambda v: "[" + ", ".join(t(x) for x in v) + "]"
    nb = Fraction(mlen) * n
    return [
        f"<r>    =   {vec(fr16[0:3])}",
        f"<r/nb> =   {vec([x / nb for x in fr16[0:3]])}",
        f"<rj2>  =   {vec(fr16[3:6])}",
        f"<r2>   =   {t(fr16[6])}",
        f"<p>    =   {vec(fr16[7:10])}",
        f"<pj2>  =   {vec(fr16[10:13])}",
        f"<p2>   =   {t(fr16[13])}",
        f"<U>    =   {t(fr16[14])}",
        f"<U2>   =   {t(fr16[15])}",
        f"AR     =   {julia_float(acc_rate)}",
    ]


def julia_vector(v) -> str:
    """Julia `show(::Vector{Float64})`: `[a, b, c]`."""
    return "[" + ", ".join(julia_float(x) for x in v) + "]"


def result_lines(avg16, acc_rate, mlen, n):
    """The 10 stdout lines, mcmc_eap_chain.jl:386-395.  avg16 is in rolling.csv column order."""
    r, rj2, r2 = avg16[0:3], avg16[3:6], avg16[6]
    p, pj2, p2 = avg16[7:10], avg16[10:13], avg16[13]
    U, U2 = avg16[14], avg16[15]
    nb = mlen * n
    return [
        f"<r>    =   {julia_vector(r)}",
        f"<r/nb> =   {julia_vector([x / nb for x in r])}",
        f"<rj2>  =   {julia_vector(rj2)}",
        f"<r2>   =   {julia_float(r2)}",
        f"<p>    =   {julia_vector(p)}",
        f"<pj2>  =   {julia_vector(pj2)}",
        f"<p2>   =   {julia_float(p2)}",
        f"<U>    =   {julia_float(U)}",
        f"<U2>   =   {julia_float(U2)}",
        f"AR     =   {julia_float(acc_rate)}",
    ]


def result_lines_2d(avg16, acc_rate, mlen, n):
    """The 10 stdout lines of the 2-D driver, 2D/mcmc_clustering_eap_chain.jl:338-347: the same as result_lines
    with 2-vectors — the planar chain lives in the x–z plane of the 3-D rows (columns "1" and "3")."""
    xz = [0, 2]
    pick = lambda v: [v[k] for k in xz]
    nb = mlen * n
    return [
        f"<r>    =   {julia_vector(pick(avg16[0:3]))}",
        f"<r/nb> =   {julia_vector([x / nb for x in pick(avg16[0:3])])}",
        f"<rj2>  =   {julia_vector(pick(avg16[3:6]))}",
        f"<r2>   =   {julia_float(avg16[6])}",
        f"<p>    =   {julia_vector(pick(avg16[7:10]))}",
        f"<pj2>  =   {julia_vector(pick(avg16[10:13]))}",
        f"<p2>   =   {julia_float(avg16[13])}",
        f"<U>    =   {julia_float(avg16[14])}",
        f"<U2>   =   {julia_float(avg16[15])}",
        f"AR     =   {julia_float(acc_rate)}",
    ]


def write_rows(fh, rows):
    """`writedlm(file, hcat(...), ',')` of Float64 rows (step is promoted to Float64, :331,:336)."""
    for row in rows:
        fh.write(",".join(julia_float(x) for x in row))
        fh.write("\n")


# ---- mcmc_clustering_eap_chain.jl: 12 stdout lines (:389-400), wider CSVs (:254-259) -------------------
ROLL_HEADER_CLUSTERING = ROLL_HEADER + ",Ealign,psi"                                       # :259


def traj_header_clustering(n: int) -> str:
    """step,r1..3,p1..3,U, phi1,theta1,…,phin,thetan, mux1,muy1,muz1,…,muzn (:254-257)."""
    cols = [TRAJ_HEADER]
    cols += [f"phi{i},theta{i}" for i in range(1, n + 1)]
    cols += [f"mux{i},muy{i},muz{i}" for i in range(1, n + 1)]
    return ",".join(cols)


def result_lines_clustering(avg16, cos2, psi, acc_rate, mlen, n):
    """The 12 stdout lines, mcmc_clustering_eap_chain.jl:389-400: the 9 averages of the plain driver, then
    `<cos2(θ)>`, `<ψ>` and AR last.  scripts/aggregate_mcmc.jl:54 names these 22 output columns."""
    base = result_lines(avg16, acc_rate, mlen, n)
    return base[:9] + [
        f"<cos2(θ)>   =   {julia_float(cos2)}",
        f"<ψ>    =   {julia_float(psi)}",
        base[9],
    ]
