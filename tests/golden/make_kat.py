"""Generates kat_energy.json: known-answer vectors for energies / observables / ΔU from an
INDEPENDENT numpy restatement of the reference formulas (vectorised, different summation order and
code from oracle/polymc_oracle.c), so that oracle ≡ numpy is a two-implementations-agree check
(SURVEY.md §8c P1).  The block "survey_n5" holds the values quoted in SURVEY.md §8c verbatim.

    python tests/golden/make_kat.py
"""
import json
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_energy.json")


def nhat(phi, th):
    return np.stack([np.cos(phi) * np.sin(th), np.sin(phi) * np.sin(th), np.cos(th)], axis=1)


def mus(nh, th, *, chain_type, E0, K1, K2, mu):
    if chain_type == "dielectric":  # dipole_response.jl:7-11
        return (K1 - K2) * E0 * np.cos(th)[:, None] * nh + K2 * np.array([0.0, 0.0, E0])
    return mu * nh                  # dipole_response.jl:27-29, M = mu I


def positions(nh, b):
    return b * (np.cumsum(nh, axis=0) - 0.5 * nh)  # eap_chain.jl:49-51


def pair_matrix(x, m):
    n = len(x)
    i, j = np.triu_indices(n, 1)
    r = x[i] - x[j]
    r2 = (r * r).sum(1)
    rm = np.sqrt(r2)
    rh = r / rm[:, None]
    e = ((m[i] * m[j]).sum(1) - 3 * (m[i] * rh).sum(1) * (m[j] * rh).sum(1)) / (4 * np.pi * r2 * rm)
    return i, j, e


def energies(phi, th, *, b, E0, Fx, Fz, chain_type, K1=1.0, K2=0.0, mu=1e-2):
    nh = nhat(phi, th)
    m = mus(nh, th, chain_type=chain_type, E0=E0, K1=K1, K2=K2, mu=mu)
    x = positions(nh, b)
    r = x[-1] + 0.5 * b * nh[-1]
    su = (-0.5 * E0 * m[:, 2]).sum()
    i, j, e = pair_matrix(x, m)
    base = su - (r[0] * Fx + r[2] * Fz)
    return {"r": r.tolist(), "p": m.sum(0).tolist(), "Omega": float(np.log(np.sin(th)).sum()),
            "su": float(su), "U_ni": float(base), "U_int": float(base + e.sum()),
            "U_ising": float(base + e[j == i + 1].sum()), "abs_pairs": float(np.abs(e).sum())}


def main():
    out = {"survey_n5": {
        "phi": [0.1, 1.3, 2.9, 4.4, 5.9], "theta": [0.4, 1.1, 1.7, 2.3, 2.9],
        "b": 1.5, "E0": 2.0, "Fx": 0.3, "Fz": 0.7,
        "r": [-0.5164145668345789, 0.5036968249916115, -0.5871323479447155], "Omega": -2.7903233230265996,
        "dielectric": {"K1": 1.0, "K2": 0.25, "p": [0.7894869500109547, 1.431610299912116, 6.1860807921389664],
                       "U_ni": -5.620163778527292, "U_int": -5.804047873024686, "U_ising": -5.8880587171766114},
        "polar": {"mu": 0.5, "p": [-0.1721381889448596, 0.16789894166387048, -0.1957107826482385],
                  "U_ni": 0.761627796259913, "U_int": 0.6695479417432036, "U_ising": 0.6826726197209616}}}
    rng = np.random.default_rng(20260101)
    cases = []
    for n in (2, 7, 33, 128):
        for chain_type, extra in (("dielectric", dict(K1=1.3, K2=0.4)), ("polar", dict(mu=0.7))):
            phi = rng.uniform(0, 2 * np.pi, n)
            th = rng.uniform(0.05, np.pi - 0.05, n)
            par = dict(b=float(rng.uniform(0.5, 2)), E0=float(rng.uniform(0.2, 3)), Fx=float(rng.uniform(-1, 1)),
                       Fz=float(rng.uniform(-2, 2)), chain_type=chain_type, **extra)
            e0 = energies(phi, th, **par)
            moves = []
            for idx in sorted({0, n - 1, n // 2, int(rng.integers(n))}):
                dphi, dth = float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-0.6, 0.6))
                phi2, th2 = phi.copy(), th.copy()
                phi2[idx] += dphi
                th2[idx] = min(np.pi, max(0.0, th2[idx] + dth))
                e1 = energies(phi2, th2, **par)
                moves.append({"idx0": idx, "dphi": dphi, "dtheta": dth,
                              "dU_ni": e1["U_ni"] - e0["U_ni"], "dU_int": e1["U_int"] - e0["U_int"],
                              "dU_ising": e1["U_ising"] - e0["U_ising"], "dOmega": e1["Omega"] - e0["Omega"],
                              "scale": e0["abs_pairs"] + e1["abs_pairs"] + abs(e0["U_ni"]) + 1.0})
            cases.append({"n": n, "phi": phi.tolist(), "theta": th.tolist(), "par": par, "E": e0, "moves": moves})
    out["random"] = cases
    json.dump(out, open(OUT, "w"), indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
