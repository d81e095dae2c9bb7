"""Generates kat_cluster.json: known-answer vectors for the clustering driver's pieces
(mcmc_clustering_eap_chain.jl; inc/eap_chain.jl:45-58 ψ/ubend, :165-192 UCutoff, :263-333 refl_n!/
cluster_flip!) from an INDEPENDENT numpy restatement (vectorised full recomputes; no changed-term
bookkeeping), so that oracle ≡ numpy is a two-implementations-agree check like kat_energy.json.

    python tests/golden/make_kat_cluster.py
"""
import json
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_cluster.json")


def nhat(phi, th):
    return np.stack([np.cos(phi) * np.sin(th), np.sin(phi) * np.sin(th), np.cos(th)], axis=1)


def mus(nh, th, *, chain_type, E0, K1, K2, mu):
    if chain_type == "dielectric":  # dipole_response.jl:7-11
        return (K1 - K2) * E0 * np.cos(th)[:, None] * nh + K2 * np.array([0.0, 0.0, E0])
    return mu * nh                  # dipole_response.jl:27-29


def everything(phi, th, *, b, E0, Fx, Fz, chain_type, K1=1.0, K2=0.0, mu=1e-2, kappa=0.0, psi0=0.0,
               cutoff_radius=7.5):
    n = len(phi)
    nh = nhat(phi, th)
    m = mus(nh, th, chain_type=chain_type, E0=E0, K1=K1, K2=K2, mu=mu)
    x = b * (np.cumsum(nh, axis=0) - 0.5 * nh)           # eap_chain.jl:49-51
    r = x[-1] + 0.5 * b * nh[-1]                         # :405
    dots = np.clip((nh[:-1] * nh[1:]).sum(1), -1.0, 1.0)
    psi = np.arccos(dots)                                # :45-47
    ubend = 0.5 * kappa * (psi - psi0) ** 2              # :54-58
    su = (-0.5 * E0 * m[:, 2]).sum() + ubend.sum()       # sum(chain.us), :130
    i, j = np.triu_indices(n, 1)
    d = x[i] - x[j]
    r2 = (d * d).sum(1)
    rm = np.sqrt(r2)
    rh = d / rm[:, None]
    e = ((m[i] * m[j]).sum(1) - 3 * (m[i] * rh).sum(1) * (m[j] * rh).sum(1)) / (4 * np.pi * r2 * rm)
    rF = r[0] * Fx + r[2] * Fz
    crad2 = (cutoff_radius * b) ** 2
    e_cut = np.where(r2 > crad2, 0.0, e)                 # :176-187
    return {
        "U_ni": float(su - rF), "U_int": float(su + e.sum() - rF), "U_ising": float(su + e[j == i + 1].sum() - rF),
        "U_cut_bare": float(e_cut.sum()), "U_cut_full": float(su + e_cut.sum() - rF),
        "Ubend": float(ubend.sum()), "psi_mean": float(psi.sum() / (n - 1)), "cos2": float((np.cos(th) ** 2).sum()),
        "Omega": float(np.log(np.sin(th)).sum()), "su": float(su), "r": r.tolist(), "p": m.sum(0).tolist(),
        "abs_pairs": float(np.abs(e).sum()), "n_cut_pairs": int((r2 <= crad2).sum()),
        "link": ((1.0 + (nh[:-1] * nh[1:]).sum(1)) / 2.0).tolist(),  # pflip_linear per bond, :267
    }


def composite(phi, th, idx, dphi, dth, reflect, lo, hi):
    """move!(idx) then refl_n! on lo..hi: the angles afterwards (eap_chain.jl:232-236, :263-265)."""
    phi2, th2 = phi.copy(), th.copy()
    phi2[idx] += dphi
    th2[idx] = min(np.pi, max(0.0, th2[idx] + dth))
    th_mid = th2.copy()
    if reflect:
        for i in range(lo, hi + 1):
            th2[i] = min(np.pi, max(0.0, th2[i] + (np.pi - 2 * th2[i])))
    return phi2, th_mid, th2


def main():
    rng = np.random.default_rng(20260102)
    cases = []
    for n in (2, 9, 40, 96):
        for chain_type, extra in (("dielectric", dict(K1=1.3, K2=0.4)), ("polar", dict(mu=0.7))):
            phi = rng.uniform(0, 2 * np.pi, n)
            th = rng.uniform(0.05, np.pi - 0.05, n)
            par = dict(b=float(rng.uniform(0.7, 1.5)), E0=float(rng.uniform(0.2, 3)), Fx=float(rng.uniform(-1, 1)),
                       Fz=float(rng.uniform(-2, 2)), kappa=float(rng.uniform(0.1, 2)), psi0=float(rng.uniform(0, 1)),
                       cutoff_radius=float(rng.uniform(1.5, 4.0)), chain_type=chain_type, **extra)
            e0 = everything(phi, th, **par)
            trials = []
            segs = {(0, 0, 0, 1), (n - 1, max(0, n - 3), n - 1, 1), (n // 2, n // 2, n // 2, 0), (n // 2, 0, n - 1, 1),
                    (n // 2, n // 2, n // 2, 1)}
            if n > 8:
                segs |= {(5, 3, 8, 1), (4, 4, n - 2, 1)}
            for (idx, lo, hi, refl) in sorted(segs):
                dphi, dth = float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-0.6, 0.6))
                phi2, th_mid, th2 = composite(phi, th, idx, dphi, dth, refl, lo, hi)
                e_mid = everything(phi2, th_mid, **par)   # the chain cluster_flip! sees (after move!)
                e1 = everything(phi2, th2, **par)
                up = e_mid["link"][hi] if hi < n - 1 else 0.0
                lp = e_mid["link"][lo - 1] if lo > 0 else 0.0
                nup = e1["link"][hi] if hi < n - 1 else 0.0
                nlp = e1["link"][lo - 1] if lo > 0 else 0.0
                alpha = ((1 - nup) * (1 - nlp)) / ((1 - up) * (1 - lp)) if refl else 1.0  # :327-330
                t = {"idx0": idx, "lo0": lo, "hi0": hi, "reflect": refl, "dphi": dphi, "dtheta": dth,
                     "log_alpha": float(np.log(alpha)),
                     "scale": e0["abs_pairs"] + e1["abs_pairs"] + abs(e0["U_ni"]) + 1.0}
                for k in ("U_ni", "U_int", "U_ising", "U_cut_bare", "U_cut_full", "Ubend", "cos2", "Omega", "su"):
                    t["d" + k] = e1[k] - e0[k]
                t["dpsi_sum"] = (e1["psi_mean"] - e0["psi_mean"]) * (n - 1)
                t["dp"] = (np.array(e1["p"]) - np.array(e0["p"])).tolist()
                t["dr"] = (np.array(e1["r"]) - np.array(e0["r"])).tolist()
                t["cut_pairs_changed"] = e1["n_cut_pairs"] != e0["n_cut_pairs"]
                trials.append(t)
            e0.pop("link")
            cases.append({"n": n, "phi": phi.tolist(), "theta": th.tolist(), "par": par, "E": e0, "trials": trials})
    json.dump({"cases": cases}, open(OUT, "w"), indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
