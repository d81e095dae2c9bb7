"""CPU tests of the oracle's restatement of the clustering driver (mcmc_clustering_eap_chain.jl):
pinned against an independent numpy restatement (tests/golden/kat_cluster.json), its own two
formulations (literal move!/refl_n! with full recomputes vs the changed-term segment ΔU), and closed
forms.  PARITY UNPINNED by the reference, as for the plain driver (no upstream tests, no Julia here)."""
import json
import math
import os

import numpy as np
import pytest

import closed_form as CF

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ET = {"U_ni": "noninteracting", "U_int": "interacting", "U_ising": "Ising", "U_cut_bare": "cutoff",
      "U_cut_full": "cutoff"}


@pytest.fixture(scope="module")
def katc():
    with open(os.path.join(GOLDEN, "kat_cluster.json")) as f:
        return json.load(f)["cases"]


def _case(O, case, key):
    return O.make_case(n=case["n"], energy_type=ET[key], clustering=True, cutoff_full=(key == "U_cut_full"),
                       **case["par"])


def test_kat_energies_with_bending_and_cutoff(O, katc):
    for case in katc:
        E = case["E"]
        scale = E["abs_pairs"] + abs(E["U_ni"]) + 1.0
        for key in ET:
            ch = O.Chain(_case(O, case, key), case["phi"], case["theta"])
            e = ch.energy_ex()
            assert abs(e["U"] - E[key]) <= 1e-12 * scale, (case["n"], key)
            assert e["su"] == pytest.approx(E["su"], rel=1e-12, abs=1e-12)
            assert e["Ubend"] == pytest.approx(E["Ubend"], rel=1e-12, abs=1e-13)
            assert e["psi"] == pytest.approx(E["psi_mean"], rel=1e-12)
            assert e["cos2"] == pytest.approx(E["cos2"], rel=1e-12)
            assert e["Omega"] == pytest.approx(E["Omega"], rel=1e-12)


def test_kat_composite_trials(O, katc):
    """move! + refl_n! on a cluster: the changed-term ΔU, the literal sequence of full recomputes and
    the independent numpy restatement agree; so does α of cluster_flip! (eap_chain.jl:317-330)."""
    for case in katc:
        n = case["n"]
        for key in ET:
            ch = O.Chain(_case(O, case, key), case["phi"], case["theta"])
            for t in case["trials"]:
                if key.startswith("U_cut") and t["cut_pairs_changed"]:
                    tol = 1e-9    # a pair sitting within rounding of the cut-off radius may flip sides
                else:
                    tol = 1e-11
                d = ch.delta_segment(t["idx0"], t["dphi"], t["dtheta"], t["reflect"], t["lo0"], t["hi0"])
                assert abs(d["dU"] - t["d" + key]) <= tol * t["scale"], (n, key, t["idx0"], t["lo0"], t["hi0"])
                if case["theta"][t["idx0"]] + t["dtheta"] <= 0.0 and t["reflect"]:
                    # move! clamps θ to 0 (Ω += −Inf), refl_n! then adds log(sin π / 0) = +Inf: the reference's
                    # running Ω is NaN and the trial is rejected; final − initial (numpy) is finite
                    assert math.isnan(d["dOmega"])
                elif math.isfinite(t["dOmega"]):
                    assert d["dOmega"] == pytest.approx(t["dOmega"], rel=1e-10, abs=1e-11)
                if math.isfinite(t["dOmega"]):
                    assert d["du"] + d["dbend"] == pytest.approx(t["dsu"], rel=1e-10, abs=1e-11)
                    assert d["dbend"] == pytest.approx(t["dUbend"], rel=1e-10, abs=1e-11)
                    assert d["dpsi"] == pytest.approx(t["dpsi_sum"], rel=1e-10, abs=1e-11)
                    assert d["dcos2"] == pytest.approx(t["dcos2"], rel=1e-10, abs=1e-11)
                    np.testing.assert_allclose([d["dp1"], d["dp2"], d["dp3"]], t["dp"], rtol=1e-10, atol=1e-11)
                # the literal form: move!, then refl_n! per cluster monomer, each a full recompute
                c2 = ch.copy()
                u0 = c2.energy_ex()["U"]
                c2.move_segment(t["idx0"], t["dphi"], t["dtheta"], t["reflect"], t["lo0"], t["hi0"])
                assert abs((c2.energy_ex()["U"] - u0) - d["dU"]) <= tol * t["scale"]
                # α from the link probabilities before / after the flip
                if t["reflect"]:
                    c3 = ch.copy()
                    c3.move(t["idx0"], t["dphi"], t["dtheta"])
                    up = c3.link_prob(t["hi0"]) if t["hi0"] < n - 1 else 0.0
                    lp = c3.link_prob(t["lo0"] - 1) if t["lo0"] > 0 else 0.0
                    nup = c2.link_prob(t["hi0"]) if t["hi0"] < n - 1 else 0.0
                    nlp = c2.link_prob(t["lo0"] - 1) if t["lo0"] > 0 else 0.0
                    la = math.log(((1 - nup) * (1 - nlp)) / ((1 - up) * (1 - lp)))
                    assert la == pytest.approx(t["log_alpha"], rel=1e-9, abs=1e-10)


def test_ucutoff_is_the_bare_pair_sum(O):
    """The reference's UCutoff functor (eap_chain.jl:171-192) returns only the pair sum — no Σu, no −r·F
    (compare InteractingEnergy, energy.jl:13-16).  Reproduced by default; cutoff_full adds them."""
    kw = dict(n=12, E0=1.5, K2=0.3, Fz=2.0, Fx=0.4, kappa=0.7, cutoff_radius=100.0, clustering=True)
    ch_i = O.Chain(O.make_case(energy_type="interacting", **kw), seed=3)
    phi, th = ch_i.state()
    ch_b = O.Chain(O.make_case(energy_type="cutoff", **kw), phi, th)
    ch_f = O.Chain(O.make_case(energy_type="cutoff", cutoff_full=True, **kw), phi, th)
    ei, eb, ef = ch_i.energy_ex(), ch_b.energy_ex(), ch_f.energy_ex()
    assert eb["U"] == pytest.approx(ei["Udd"], rel=1e-13)           # huge radius: every pair is inside
    assert ef["U"] == pytest.approx(ei["U"], rel=1e-13)
    assert abs(eb["U"] - ef["U"]) > 1.0


@pytest.mark.parametrize("et,n,steps", [("noninteracting", 40, 3000), ("Ising", 40, 3000), ("interacting", 30, 1200),
                                        ("cutoff", 30, 1200)])
@pytest.mark.parametrize("umbrella,carry", [(False, True), (True, False)])
def test_two_formulations_same_trajectory(O, et, n, steps, umbrella, carry):
    """algo 0 = the reference's own sequence (deep copy, move!, refl_n! with full recomputes, stateful
    acceptor with α) and algo 1 = the changed-term form take the same decisions, through a kT ladder."""
    c = O.make_case(n=n, energy_type=et, E0=1.0, K2=0.2, Fz=0.5, Fx=0.1, kappa=0.5, psi0=0.2, cutoff_radius=3.0,
                    clustering=True, alpha_carry=carry, umbrella=umbrella, adj_ub=0.4, steps_per_adjust=200)
    r0, r1 = O.Run(c, 42, 7, 0), O.Run(c, 42, 7, 1)
    for kT in (10.0, 1.0):
        r0.begin_stage(kT)
        r1.begin_stage(kT)
        t0, l0, s0 = r0.steps_ex(steps, 100, True)
        t1, l1, s1 = r1.steps_ex(steps, 100, True)
        np.testing.assert_array_equal(s0, s1)
        np.testing.assert_allclose(t0, t1, rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(l0, l1, rtol=1e-8, atol=1e-7)
        assert r0.diag()["nacc_total"] == r1.diag()["nacc_total"]
        assert r0.cluster_stats() == r1.cluster_stats()
        np.testing.assert_allclose(r0.extra_averages(), r1.extra_averages(), rtol=1e-10)
    assert l0.shape[1] == 19 and t0[0, 0] == 100.0


def test_stage_restarts_like_a_fresh_mcmc_call(O):
    """mcmc(nsteps, pargs, chain) (:171-265): averagers, counters and step sizes restart, the chain stays."""
    c = O.make_case(n=20, energy_type="Ising", E0=1.0, Fz=0.3, clustering=True, phi_step=0.01, theta_step=0.005,
                    steps_per_adjust=50, adj_scale=2.0, adj_ub=0.4)
    r = O.Run(c, 5, 0, 1)
    r.begin_stage(100.0)
    r.steps_ex(200, 0)
    d = r.diag()
    assert d["phi_step"] > 0.01 and d["steps_total"] == 200
    phi, th = r.chain().state()
    r.begin_stage(1.0)
    d = r.diag()
    assert d["phi_step"] == 0.01 and d["theta_step"] == 0.005 and d["steps_total"] == 0 and d["nacc_total"] == 0
    phi2, th2 = r.chain().state()
    np.testing.assert_array_equal(phi, phi2)
    np.testing.assert_array_equal(th, th2)
    r.steps_ex(300, 0)
    assert r.averages()[2] == 300


def test_cluster_gate_and_growth_statistics(O):
    """cluster_flip! returns early iff rand() <= cluster-prob (:273), so 1−cluster-prob of the trials flip a
    cluster; clusters always contain idx; cluster-prob = 1 never flips."""
    c = O.make_case(n=50, energy_type="noninteracting", clustering=True, cluster_prob=0.25, E0=0.0)
    r = O.Run(c, 11, 0, 1)
    r.begin_stage(1.0)
    r.steps_ex(8000, 0)
    s = r.cluster_stats()
    assert s["ncluster"] / 8000 == pytest.approx(0.75, abs=0.02)
    assert s["cluster_sum"] / s["ncluster"] > 1.5 and s["cluster_max"] <= 50
    c = O.make_case(n=50, energy_type="noninteracting", clustering=True, cluster_prob=1.0, E0=0.0)
    r = O.Run(c, 11, 0, 1)
    r.begin_stage(1.0)
    r.steps_ex(2000, 0)
    assert r.cluster_stats()["ncluster"] == 0


def test_x0_initial_configuration(O):
    """--x0/--dx0 (eap_chain.jl:63-78): ϕ = ϕ0 + U(0,dx0[1]), θ = θ0 + U(0,dx0[2])."""
    c = O.make_case(n=16, energy_type="Ising", clustering=True)
    r = O.Run(c, 9, 2, 1)
    r.init_x0([0.0, math.pi / 2], [2 * math.pi, 0.1])
    phi, th = r.chain().state()
    assert np.all((phi >= 0) & (phi < 2 * math.pi)) and np.all((th >= math.pi / 2) & (th < math.pi / 2 + 0.1))
    x0 = np.arange(32, dtype=float) * 0.01
    r.init_x0(x0, [0.0, 0.0])
    phi, th = r.chain().state()
    np.testing.assert_array_equal(phi, x0[0::2])
    np.testing.assert_array_equal(th, x0[1::2])


def test_bending_only_chain_matches_bond_angle_closed_form(O):
    """E0 = 0, F = 0, κ > 0, no cluster flips: bond angles are independent with density
    ∝ sinψ exp(−κ(ψ−ψ0)²/2kT); ⟨Σψ/(n−1)⟩ against the quadrature, 3.5σ over independent chains."""
    kappa, psi0 = 2.0, 0.4
    want = CF.bond_angle_mean(kappa, psi0)
    c = O.make_case(n=12, energy_type="noninteracting", clustering=True, cluster_prob=1.0, kappa=kappa, psi0=psi0,
                    adj_ub=0.4)
    vals = []
    for cid in range(12):
        r = O.Run(c, 2024, cid, 1)
        r.begin_stage(1.0)
        r.steps_ex(5000, 0)
        r.begin_stage(1.0)
        r.steps_ex(60000, 0)
        vals.append(r.extra_averages()[1])
    v = np.array(vals)
    sem = v.std(ddof=1) / math.sqrt(len(v))
    assert abs(v.mean() - want) <= 3.5 * sem, (v.mean(), want, sem)


def test_reference_acceptor_keeps_log_alpha(O):
    """inc/acceptance.jl:30-33 stores logπ + log α as logπ_prev.  The oracle reproduces it by default
    (alpha_carry); the two settings are different Markov chains (different decisions on the same stream)."""
    kw = dict(n=30, energy_type="noninteracting", clustering=True, E0=2.0, Fz=1.5, adj_ub=0.4)
    a = O.Run(O.make_case(alpha_carry=True, **kw), 99, 0, 1)
    b = O.Run(O.make_case(alpha_carry=False, **kw), 99, 0, 1)
    for r in (a, b):
        r.begin_stage(1.0)
        r.steps_ex(20000, 0)
    assert a.diag()["nacc_total"] != b.diag()["nacc_total"]
