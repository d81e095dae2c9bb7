"""Generates kat_2d.json: known-answer vectors for the 2-D tree (2D/inc/eap_chain.jl, dipole_response.jl,
energy.jl) from an INDEPENDENT numpy restatement with genuine 2-vectors (the oracle and the CUDA path embed
the plane in their 3-vectors), incl. composite trials move! + flip_n! on a cluster and α.

    python tests/golden/make_kat_2d.py
"""
import json
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_2d.json")


def everything(phi, *, b, E0, Fx, Fz, chain_type, K1=1.0, K2=0.0, mu=1e-2):
    n = len(phi)
    nh = np.stack([np.cos(phi), np.sin(phi)], axis=1)                 # 2D/inc/eap_chain.jl:33
    if chain_type == "dielectric":                                    # 2D/inc/dipole_response.jl:7-10
        m = (K1 - K2) * E0 * np.sin(phi)[:, None] * nh + K2 * np.array([0.0, E0])
    else:                                                             # :25-27
        m = mu * nh
    x = b * (np.cumsum(nh, axis=0) - 0.5 * nh)                        # 2D/inc/eap_chain.jl:49-51
    r = x[-1] + 0.5 * b * nh[-1]
    su = (-0.5 * E0 * m[:, 1]).sum()                                  # u = −½E0μ[2], :64
    i, j = np.triu_indices(n, 1)
    d = x[i] - x[j]
    r2 = (d * d).sum(1)
    rm = np.sqrt(r2)
    rh = d / rm[:, None]
    e = ((m[i] * m[j]).sum(1) - 3 * (m[i] * rh).sum(1) * (m[j] * rh).sum(1)) / (4 * np.pi * r2 * rm)   # :141-151
    rF = r[0] * Fx + r[1] * Fz                                        # 2D/inc/energy.jl:8
    return {"U_ni": float(su - rF), "U_int": float(su + e.sum() - rF), "U_ising": float(su + e[j == i + 1].sum() - rF),
            "su": float(su), "r": r.tolist(), "p": m.sum(0).tolist(), "abs_pairs": float(np.abs(e).sum()),
            "link": ((1.0 + (nh[:-1] * nh[1:]).sum(1)) / 2.0).tolist()}


def main():
    rng = np.random.default_rng(20260103)
    cases = []
    for n in (2, 9, 40, 96):
        for chain_type, extra in (("dielectric", dict(K1=1.3, K2=0.4)), ("polar", dict(mu=0.7))):
            phi = rng.uniform(0, 2 * np.pi, n)
            par = dict(b=float(rng.uniform(0.7, 1.5)), E0=float(rng.uniform(0.2, 3)), Fx=float(rng.uniform(-1, 1)),
                       Fz=float(rng.uniform(-2, 2)), chain_type=chain_type, **extra)
            e0 = everything(phi, **par)
            trials = []
            segs = {(0, 0, 0, 1), (n - 1, max(0, n - 3), n - 1, 1), (n // 2, n // 2, n // 2, 0), (n // 2, 0, n - 1, 1)}
            if n > 8:
                segs |= {(5, 3, 8, 1), (4, 4, n - 2, 1)}
            for (idx, lo, hi, refl) in sorted(segs):
                dphi = float(rng.uniform(-1.2, 1.2))
                phi_mid = phi.copy()
                phi_mid[idx] += dphi                                  # move!, 2D/inc/eap_chain.jl:171-187
                phi2 = phi_mid.copy()
                if refl:
                    phi2[lo:hi + 1] += np.pi                          # flip_n!, :189-191
                e_mid, e1 = everything(phi_mid, **par), everything(phi2, **par)
                up = e_mid["link"][hi] if hi < n - 1 else 0.0
                lp = e_mid["link"][lo - 1] if lo > 0 else 0.0
                nup = e1["link"][hi] if hi < n - 1 else 0.0
                nlp = e1["link"][lo - 1] if lo > 0 else 0.0
                alpha = ((1 - nup) * (1 - nlp)) / ((1 - up) * (1 - lp)) if refl else 1.0   # :246-249
                t = {"idx0": idx, "lo0": lo, "hi0": hi, "reflect": refl, "dphi": dphi, "log_alpha": float(np.log(alpha)),
                     "scale": e0["abs_pairs"] + e1["abs_pairs"] + abs(e0["U_ni"]) + 1.0}
                for k in ("U_ni", "U_int", "U_ising", "su"):
                    t["d" + k] = e1[k] - e0[k]
                t["dp"] = (np.array(e1["p"]) - np.array(e0["p"])).tolist()
                t["dr"] = (np.array(e1["r"]) - np.array(e0["r"])).tolist()
                trials.append(t)
            e0.pop("link")
            cases.append({"n": n, "phi": phi.tolist(), "par": par, "E": e0, "trials": trials})
    json.dump({"cases": cases}, open(OUT, "w"), indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
