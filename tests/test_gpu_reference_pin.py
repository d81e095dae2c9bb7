"""The CUDA library against outputs of the reference itself (tests/golden/ref/fine_seams.json — the unmodified
inc/eap_chain.jl, inc/energy.jl, inc/acceptance.jl executed on fixed chains, see tests/golden/make_ref_fixtures.py):
U(chain) of every energy functor, Ω, r, p, ψ, `move!` as ΔU, `cluster_flip!` as the segment update with its α — through
the C ABI, to 1e-12 relative (normalised by Σ|pair terms|, SURVEY finding 8).  No oracle in between.
"""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


@pytest.fixture(scope="module")
def fine():
    with open(os.path.join(ROOT, "tests", "golden", "ref", "fine_seams.json")) as f:
        return json.load(f)


def num(x):
    return float(x) if not isinstance(x, str) else {"NaN": math.nan, "Inf": math.inf, "-Inf": -math.inf}[x]


def gpu_case(pm, c, **over):
    kw = dict(n=c["n"], E0=c["E0"], K1=c["K1"], K2=c["K2"], mu=c["mu"], kT=c["kT"], Fz=c["Fz"], Fx=c["Fx"], b=c["b"],
              chain_type=c["chain_type"], energy_type=c["energy_type"], kappa=c["kappa"], psi0=c["psi0"],
              cutoff_radius=c["cutoff_radius"], clustering=True)
    kw.update(over)
    return pm.make_case(**kw)


def pair_scale(c):
    """Σ|pair terms| is not in the fixture; |U_interaction| + |U| + Σ|u| bounds the size of what was summed from below,
    and the all-pairs sum of a random chain is dominated by its closest contact — use the largest of the three sums."""
    return 1.0 + abs(c["U"]) + abs(c["sum_us"]) + abs(c["U_interaction"]) + abs(c["U_cutoff"]) + abs(c["U_Ising"])


def test_energies_match_the_reference_on_gpu(pm, fine):
    for c in fine["cases"]:
        with pm.Ensemble(gpu_case(pm, c), replicas=2, seed=1) as ens:
            ens.set_state(1, c["phi"], c["theta"])
            e = ens.energy_ex(1)
            sc = pair_scale(c)
            assert abs(e["U"] - c["U"]) <= TOL * sc, (c["name"], e["U"], c["U"])
            assert abs(e["su"] - c["sum_us"]) <= TOL * sc
            assert e["Omega"] == pytest.approx(c["Omega"], rel=1e-12, abs=1e-12)
            assert e["psi"] * (c["n"] - 1) == pytest.approx(c["sum_psi"], rel=1e-12)
            assert e["cos2"] == pytest.approx(c["sum_cos2"], rel=1e-12)
            r, p = ens.observables(1)
            np.testing.assert_allclose(r, c["r"], rtol=1e-12, atol=1e-12 * c["n"])
            np.testing.assert_allclose(p, c["p"], rtol=1e-12, atol=1e-12 * c["n"])
        for et, key in (("interacting", "U_interaction"), ("Ising", "U_Ising"), ("cutoff", "U_cutoff")):
            with pm.Ensemble(gpu_case(pm, c, energy_type=et), replicas=1, seed=1) as ens:
                ens.set_state(0, c["phi"], c["theta"])
                assert abs(ens.energy(0)["Udd"] - c[key]) <= TOL * pair_scale(c), (c["name"], key)
        # the plain driver's kernels (no bending, no clustering flag) on the same chains
        if c["kappa"] == 0.0 and c["energy_type"] != "cutoff":
            with pm.Ensemble(gpu_case(pm, c, clustering=False), replicas=1, seed=1) as ens:
                ens.set_state(0, c["phi"], c["theta"])
                assert abs(ens.energy(0)["U"] - c["U"]) <= TOL * pair_scale(c)


def test_moves_match_the_reference_on_gpu(pm, fine):
    """pmc_delta_u / pmc_delta_segment ≡ U(move!(copy)) − U(chain) of the reference."""
    for c in fine["cases"]:
        handles = [pm.Ensemble(gpu_case(pm, c), replicas=1, seed=1)]
        if c["kappa"] == 0.0 and c["energy_type"] != "cutoff":
            handles.append(pm.Ensemble(gpu_case(pm, c, clustering=False), replicas=1, seed=1))   # single-monomer kernels
        try:
            for ens in handles:
                ens.set_state(0, c["phi"], c["theta"])
                for m in c["moves"]:
                    Uref, Oref = num(m["U"]), num(m["Omega"])
                    d = ens.delta_u(0, m["idx"] - 1, m["dphi"], m["dtheta"])
                    sc = 20 * TOL * (pair_scale(c) + abs(Uref))
                    assert abs(d["dU"] - (Uref - c["U"])) <= sc, (c["name"], m["idx"], d["dU"], Uref - c["U"])
                    if math.isfinite(Oref):
                        assert d["dOmega"] == pytest.approx(Oref - c["Omega"], rel=1e-10, abs=1e-11)
                    else:
                        assert d["dOmega"] == -math.inf and d["clamped"]
        finally:
            for ens in handles:
                ens.close()


def test_cluster_flips_match_the_reference_on_gpu(pm, fine):
    """pmc_delta_segment with the reference's own cluster bounds: ΔU, ΔΩ, Δp, ΔΣψ, ΔΣcos²θ and log α of
    `move!` + `cluster_flip!` (inc/eap_chain.jl:269-333)."""
    nflip = 0
    for c in fine["cases"]:
        with pm.Ensemble(gpu_case(pm, c), replicas=1, seed=1) as ens:
            ens.set_state(0, c["phi"], c["theta"])
            for f in c["cluster_flips"]:
                idx0 = f["idx"] - 1
                reflect = f["lo"] > 0
                lo0, hi0 = (f["lo"] - 1, f["hi"] - 1) if reflect else (idx0, idx0)
                d = ens.delta_segment(0, idx0, f["dphi"], f["dtheta"], reflect, lo0, hi0)
                Uref = num(f["U"])
                sc = 20 * TOL * (pair_scale(c) + abs(Uref))
                assert abs(d["dU"] - (Uref - c["U"])) <= sc, (c["name"], f["idx"], d["dU"], Uref - c["U"])
                if math.isfinite(num(f["Omega"])):
                    assert d["dOmega"] == pytest.approx(num(f["Omega"]) - c["Omega"], rel=1e-10, abs=1e-11)
                np.testing.assert_allclose([d["dp1"], d["dp2"], d["dp3"]], np.array(f["p"]) - np.array(c["p"]), rtol=0,
                                           atol=1e-11 * (1 + np.abs(c["p"]).max()))
                assert d["dpsi"] == pytest.approx(f["sum_psi"] - c["sum_psi"], abs=1e-11)
                assert d["dcos2"] == pytest.approx(f["sum_cos2"] - c["sum_cos2"], abs=1e-11)
                if reflect:
                    nflip += 1
                    assert math.exp(d["log_alpha"]) == pytest.approx(f["alpha"], rel=1e-11)
                else:
                    assert d["log_alpha"] == 0.0 and f["alpha"] == 1.0
    assert nflip >= 30
