"""Closed-form (quadrature) ensemble averages for NON-INTERACTING chains — pin P2 of SURVEY.md §8c.

TEST INFRASTRUCTURE ONLY (same rule as oracle.py).

With `--energy-type noninteracting` (inc/energy.jl:7-9) the chain energy is a sum of
single-monomer terms e(θ,ϕ) = u(θ) − b (Fx n̂x + Fz n̂z), u = −½ E0 μz (inc/eap_chain.jl:53),
so the monomers are independent with density ∝ sinθ · exp(−e/kT) on the sphere.  The sampler
clamps θ to [0,π] (eap_chain.jl:236), and a clamped proposal has sinθ = 0 (or 1.2e-16) and is
rejected, which is plain Metropolis on the restricted domain — the target density is unchanged.

For E0 = 0, Fx = 0 the result reduces to the Langevin function ⟨cosθ⟩ = coth f − 1/f, f = bFz/kT
(the inverse of which the reference approximates in inc/langevin.jl:1-3).
"""
from __future__ import annotations

import numpy as np


def _mu(x, sth, cph, sph, *, chain_type, E0, K1, K2, mu):
    """Dipole response, inc/dipole_response.jl:7-11 (dielectric) / :27-29 with M = mu·I (polar)."""
    nx, ny, nz = cph * sth, sph * sth, x
    if chain_type == "dielectric":
        f = (K1 - K2) * E0 * x
        return f * nx, f * ny, f * nz + K2 * E0
    return mu * nx, mu * ny, mu * nz


def single_monomer_moments(*, chain_type="dielectric", E0=0.0, K1=1.0, K2=0.0, mu=1e-2, kT=1.0,
                           Fz=0.0, Fx=0.0, b=1.0, nx_quad=400, nphi_quad=256):
    """Moments of one monomer under density ∝ exp(−e/kT) dx dϕ (x = cosθ).

    Gauss–Legendre in x, periodic trapezoid in ϕ (spectrally accurate)."""
    xs, wx = np.polynomial.legendre.leggauss(nx_quad)
    ph = np.arange(nphi_quad) * (2 * np.pi / nphi_quad)
    X, P = np.meshgrid(xs, ph, indexing="ij")
    W = wx[:, None] * np.full_like(P, 2 * np.pi / nphi_quad)
    sth = np.sqrt(np.maximum(0.0, 1 - X * X))
    cph, sph = np.cos(P), np.sin(P)
    mx, my, mz = _mu(X, sth, cph, sph, chain_type=chain_type, E0=E0, K1=K1, K2=K2, mu=mu)
    nx, ny, nz = cph * sth, sph * sth, X
    e = -0.5 * E0 * mz - b * (Fx * nx + Fz * nz)
    logw = -e / kT
    logw -= logw.max()
    dens = np.exp(logw) * W
    Z = dens.sum()

    def avg(f):
        return float((f * dens).sum() / Z)

    return {
        "n": np.array([avg(nx), avg(ny), avg(nz)]),
        "n2": np.array([avg(nx * nx), avg(ny * ny), avg(nz * nz)]),
        "mu": np.array([avg(mx), avg(my), avg(mz)]),
        "mu2": np.array([avg(mx * mx), avg(my * my), avg(mz * mz)]),
        "e": avg(e),
        "e2": avg(e * e),
    }


def chain_averages(n, **kw):
    """The 16 averages in rolling.csv order (mcmc_eap_chain.jl:259) for n independent monomers."""
    b = kw.get("b", 1.0)
    m = single_monomer_moments(**kw)
    r = n * b * m["n"]
    rj2 = b * b * (n * m["n2"] + n * (n - 1) * m["n"] ** 2)
    p = n * m["mu"]
    pj2 = n * m["mu2"] + n * (n - 1) * m["mu"] ** 2
    U = n * m["e"]
    U2 = n * (m["e2"] - m["e"] ** 2) + (n * m["e"]) ** 2
    return np.concatenate([r, rj2, [rj2.sum()], p, pj2, [pj2.sum()], [U, U2]])


def langevin(f):
    """⟨cosθ⟩ of a freely-jointed monomer under reduced force f = bF/kT: coth f − 1/f."""
    f = np.asarray(f, dtype=float)
    small = np.abs(f) < 1e-4
    fs = np.where(small, 1.0, f)
    return np.where(small, f / 3 - f ** 3 / 45, 1 / np.tanh(fs) - 1 / fs)


def bond_angle_mean(kappa, psi0=0.0, kT=1.0, nquad=2000):
    """⟨ψ⟩ of a chain whose ONLY energy is bending, ubend = κ/2 (ψ−ψ0)² (inc/eap_chain.jl:54-58; E0 = 0,
    F = 0): the bond angles are independent with density ∝ sinψ · exp(−κ(ψ−ψ0)²/(2kT)) on [0,π]
    (each direction is uniform on the sphere given its predecessor).  The clustering driver averages
    Σψ/(n−1) (mcmc_clustering_eap_chain.jl:244)."""
    xs, w = np.polynomial.legendre.leggauss(nquad)
    psi = 0.5 * np.pi * (xs + 1.0)
    dens = np.sin(psi) * np.exp(-kappa * (psi - psi0) ** 2 / (2 * kT)) * w
    return float((psi * dens).sum() / dens.sum())


def cos2_sum(n, **kw):
    """⟨Σcos²θ⟩ of n independent monomers (mcmc_clustering_eap_chain.jl:243)."""
    return n * float(single_monomer_moments(**kw)["n2"][2])


def planar_chain_averages(n, *, chain_type="dielectric", E0=0.0, K1=1.0, K2=0.0, mu=1e-2, kT=1.0, Fz=0.0, Fx=0.0,
                          b=1.0, nquad=4096):
    """The 16 averages (3-D rolling.csv order, y components zero) of n independent PLANAR monomers
    (2D/inc/eap_chain.jl): n̂ = (cosϕ, sinϕ), field along the second axis, flat measure dϕ — density
    ∝ exp(−e(ϕ)/kT), e = −½E0μ₂ − b(Fx cosϕ + Fz sinϕ) (2D/inc/energy.jl:7-9, 2D/inc/eap_chain.jl:64)."""
    ph = np.arange(nquad) * (2 * np.pi / nquad)   # periodic trapezoid: spectrally accurate
    c, s = np.cos(ph), np.sin(ph)
    if chain_type == "dielectric":                # 2D/inc/dipole_response.jl:7-10
        f = (K1 - K2) * E0 * s
        mx, mz = f * c, f * s + K2 * E0
    else:                                         # :25-27
        mx, mz = mu * c, mu * s
    e = -0.5 * E0 * mz - b * (Fx * c + Fz * s)
    w = np.exp(-(e - e.min()) / kT)
    w /= w.sum()

    def avg(fv):
        return float((fv * w).sum())

    nbar = np.array([avg(c), 0.0, avg(s)])
    n2 = np.array([avg(c * c), 0.0, avg(s * s)])
    mbar = np.array([avg(mx), 0.0, avg(mz)])
    m2 = np.array([avg(mx * mx), 0.0, avg(mz * mz)])
    eb, e2 = avg(e), avg(e * e)
    r = n * b * nbar
    rj2 = b * b * (n * n2 + n * (n - 1) * nbar ** 2)
    p = n * mbar
    pj2 = n * m2 + n * (n - 1) * mbar ** 2
    U = n * eb
    U2 = n * (e2 - eb ** 2) + (n * eb) ** 2
    return np.concatenate([r, rj2, [rj2.sum()], p, pj2, [pj2.sum()], [U, U2]])
