"""CPU tests of the oracle: pinned against golden vectors (independent numpy restatement and the
values quoted in SURVEY.md §8c), closed forms, and its own two formulations."""
import math

import numpy as np
import pytest

import closed_form as CF

ET = {"U_ni": "noninteracting", "U_int": "interacting", "U_ising": "Ising"}


def test_philox_known_answers(O):
    # Random123 kat_vectors for philox4x32-10
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draws_are_in_range_and_distinct(O):
    seen = set()
    for step in range(1, 200):
        idx, up, flip, ut, eps = O.draw_step(9, 3, 0, step, 100)
        assert 0 <= idx < 100 and 0 <= up < 1 and 0 <= ut < 1 and 0 <= eps < 1 and flip in (0, 1)
        seen.add((idx, up))
    assert len(seen) == 199
    phi, th = O.draw_init(9, 3, 0, 64)
    assert np.all((phi >= 0) & (phi < 2 * math.pi)) and np.all((th >= 0) & (th < math.pi))
    assert O.draw_step(9, 3, 0, 5, 100) != O.draw_step(9, 4, 0, 5, 100)
    assert O.draw_step(9, 3, 0, 5, 100) != O.draw_step(9, 3, 1, 5, 100)


def test_survey_quoted_kat(O, kat):
    k = kat["survey_n5"]
    for ct in ("dielectric", "polar"):
        extra = {x: k[ct][x] for x in ("K1", "K2", "mu") if x in k[ct]}
        for key, et in ET.items():
            c = O.make_case(n=5, b=k["b"], E0=k["E0"], Fx=k["Fx"], Fz=k["Fz"], chain_type=ct, energy_type=et, **extra)
            ch = O.Chain(c, k["phi"], k["theta"])
            assert ch.energy()["U"] == pytest.approx(k[ct][key], rel=1e-13, abs=1e-13)
            assert ch.energy()["Omega"] == pytest.approx(k["Omega"], rel=1e-13)
            np.testing.assert_allclose(ch.r(), k["r"], rtol=1e-13, atol=1e-14)
            np.testing.assert_allclose(ch.p(), k[ct]["p"], rtol=1e-13, atol=1e-14)


def test_random_kat_energies_and_moves(O, kat):
    for case in kat["random"]:
        par = case["par"]
        for key, et in ET.items():
            c = O.make_case(n=case["n"], energy_type=et, **par)
            ch = O.Chain(c, case["phi"], case["theta"])
            scale = case["E"]["abs_pairs"] + abs(case["E"]["U_ni"]) + 1.0
            assert abs(ch.energy()["U"] - case["E"][key]) <= 1e-12 * scale
            assert ch.energy()["Omega"] == pytest.approx(case["E"]["Omega"], rel=1e-12)
            np.testing.assert_allclose(ch.r(), case["E"]["r"], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(ch.p(), case["E"]["p"], rtol=1e-12, atol=1e-12)
            for mv in case["moves"]:
                d = ch.delta_u(mv["idx0"], mv["dphi"], mv["dtheta"])
                want = mv["d" + key]
                assert abs(d["dU"] - want) <= 1e-11 * mv["scale"], (case["n"], et, mv["idx0"])
                if math.isfinite(mv["dOmega"]):
                    assert d["dOmega"] == pytest.approx(mv["dOmega"], rel=1e-10, abs=1e-12)
                else:
                    assert d["dOmega"] == -math.inf
                # move! (full recompute) agrees with the changed-pair ΔU
                c2 = ch.copy()
                u0 = c2.energy()["U"]
                c2.move(mv["idx0"], mv["dphi"], mv["dtheta"])
                assert abs((c2.energy()["U"] - u0) - d["dU"]) <= 1e-11 * mv["scale"]


def test_full_recompute_conditioning(O):
    """The reference's full recompute is ill-conditioned near contacts (no excluded volume, SURVEY
    finding 8): U(move!(copy)) − U(chain) carries ≈3·eps·|x|/r_min·Σ|terms| of noise, while the
    changed-pair ΔU does not.  This is why GPU-vs-recompute drift is judged against Σ|pair terms|."""
    n = 512
    oc = O.make_case(n=n, E0=1.0, Fz=0.5, energy_type="interacting")
    seen_large = False
    for c in range(40):
        ch = O.Chain(oc, seed=20260101, chain_id=c)
        S = ch.abs_pair_sum()
        d = ch.delta_u(7, 0.3, 0.2)
        c2 = ch.copy()
        c2.move(7, 0.3, 0.2)
        incons = abs((c2.energy()["U"] - ch.energy()["U"]) - d["dU"])
        assert incons <= 1e-9 * (S + 1.0)            # bounded relative to Σ|terms| (with contact amplification)
        seen_large |= incons > 1e-12 * max(1.0, d["abs_sum"])
    assert seen_large                                # ...but far above the ΔU's own rounding level


def test_changed_pair_count_formula(O):
    # SURVEY finding 2: (idx)(n-1-idx) + (n-1) changed pairs; mean (n-1)(n-2)/6 + (n-1)
    n = 512
    mean = np.mean([i * (n - 1 - i) + (n - 1) for i in range(n)])
    assert mean == pytest.approx((n - 1) * (n - 2) / 6 + (n - 1))
    assert round(mean) == 43946


@pytest.mark.parametrize("et,n,steps", [("noninteracting", 50, 4000), ("Ising", 50, 4000), ("interacting", 40, 1500)])
@pytest.mark.parametrize("flips,umbrella", [(False, False), (True, True)])
def test_two_formulations_same_trajectory(O, et, n, steps, flips, umbrella):
    c = O.make_case(n=n, energy_type=et, E0=1.0, K2=0.2, Fz=0.5, Fx=0.1, do_flips=flips, umbrella=umbrella,
                    steps_per_adjust=200)
    r0, r1 = O.Run(c, 42, 7, 0), O.Run(c, 42, 7, 1)
    t0, l0 = r0.steps(steps, 100)
    t1, l1 = r1.steps(steps, 100)
    np.testing.assert_allclose(t0, t1, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(l0, l1, rtol=1e-9, atol=1e-9)
    assert r0.diag()["nacc_total"] == r1.diag()["nacc_total"]
    assert r0.diag()["phi_step"] == r1.diag()["phi_step"]


def test_step_adaptation_rule(O):
    # high acceptance (tiny steps, no field) must grow the steps by exactly `scale`, capped at π, π/2
    c = O.make_case(n=20, phi_step=0.01, theta_step=0.005, steps_per_adjust=50, adj_scale=2.0)
    r = O.Run(c, 1, 0, 1)
    r.steps(50)
    d = r.diag()
    assert d["phi_step"] == pytest.approx(0.02) and d["theta_step"] == pytest.approx(0.01)
    assert d["nacc"] == 0 and d["natt"] == 0 and d["nacc_total"] > 25
    r.steps(2000)
    d = r.diag()
    assert d["phi_step"] <= math.pi and d["theta_step"] <= math.pi / 2
    # scale == 1 disables adaptation (mcmc_eap_chain.jl:302)
    c = O.make_case(n=20, phi_step=0.01, theta_step=0.005, steps_per_adjust=50, adj_scale=1.0)
    r = O.Run(c, 1, 0, 1)
    r.steps(500)
    assert r.diag()["phi_step"] == 0.01 and r.diag()["natt"] == 500


def test_clamped_theta_is_rejected(O):
    c = O.make_case(n=8, energy_type="interacting", E0=1.0)
    ch = O.Chain(c, seed=3)
    d = ch.delta_u(2, 0.1, -10.0)           # θ' clamps to 0 ⇒ sinθ' = 0 ⇒ ΔΩ = −Inf
    assert d["dOmega"] == -math.inf
    d = ch.delta_u(2, 0.1, +10.0)           # θ' clamps to π ⇒ sin(π) = 1.2e-16 ⇒ finite but ≈ −36
    assert math.isfinite(d["dOmega"]) and d["dOmega"] < -30


def test_omega_compat_reproduces_reference_underflow(O):
    # SURVEY finding 7: log(prod(sinθ)) underflows for n >~ 1100 ⇒ −Inf ⇒ every move accepted.
    c = O.make_case(n=1500, omega_compat=True)
    r = O.Run(c, 5, 0, 0)
    assert r.chain().energy()["Omega"] == -math.inf
    r.steps(300)
    # (a monomer accepted at θ=0 later yields −Inf+Inf = NaN, the only rejections left)
    assert r.averages()[1] > 0.98
    # the default (Σ log sinθ) is finite and samples normally
    r = O.Run(O.make_case(n=1500), 5, 0, 0)
    assert math.isfinite(r.chain().energy()["Omega"])
    r.steps(300)
    assert 0.2 < r.averages()[1] < 0.95


def test_reinit(O):
    c = O.make_case(n=30, energy_type="Ising", E0=1.0, Fz=1.0)
    r = O.Run(c, 8, 1, 1)
    r.steps(500, 100)
    phi0, _ = r.chain().state()
    took = r.reinit(force=True)
    assert took
    phi1, th1 = r.chain().state()
    ephi, eth = O.draw_init(8, 1, 1, 30)
    np.testing.assert_array_equal(phi1, ephi)
    np.testing.assert_array_equal(th1, eth)
    t, _ = r.steps(200, 100)
    assert t[0, 0] == 100.0 and r.averages()[2] == 700


def _batch_means(roll, col, discard):
    """Per-batch means from cumulative averages: batch_k = k·A_k − (k−1)·A_{k−1}."""
    k = np.arange(1, roll.shape[0] + 1)
    cum = roll[:, col] * k
    b = np.diff(np.concatenate([[0.0], cum]))
    return b[discard:]


@pytest.mark.parametrize("kw", [
    dict(E0=0.0, Fz=1.5),                       # freely jointed chain: Langevin
    dict(E0=2.0, K1=1.0, K2=0.0, Fz=1.5),       # dielectric in a field
    dict(E0=1.0, K1=0.5, K2=1.0, Fz=0.5, Fx=0.7),
    dict(E0=2.0, mu=1.5, Fz=-0.5, chain_type="polar"),
])
def test_noninteracting_matches_closed_form(O, kw):
    """P2: config C1 — n=100 non-interacting chain vs the single-monomer quadrature, 3σ batch means."""
    n = 100
    cf = CF.chain_averages(n, **kw)
    c = O.make_case(n=n, energy_type="noninteracting", **kw)
    r = O.Run(c, 1234, 0, 1)
    _, roll = r.steps(400000, 4000)
    for col, name in ((3, "r3"), (1, "r1"), (15, "U"), (10, "p3")):
        b = _batch_means(roll, col, discard=5)
        mean, sem = b.mean(), b.std(ddof=1) / math.sqrt(len(b))
        want = cf[col - 1]
        assert abs(mean - want) <= 3.5 * sem + 1e-9 * max(1, abs(want)), (name, mean, want, sem)
    if kw.get("E0") == 0.0:
        assert cf[2] / n == pytest.approx(float(CF.langevin(1.5)), rel=1e-9)


def test_umbrella_reweighting_is_consistent(O):
    """Umbrella-sampled averages (average.jl:63-97) estimate the same ensemble averages."""
    kw = dict(E0=1.5, K1=1.0, K2=0.0, Fz=0.3)
    n = 20
    cf = CF.chain_averages(n, **kw)
    c = O.make_case(n=n, energy_type="noninteracting", umbrella=True, **kw)
    r = O.Run(c, 77, 0, 1)
    _, roll = r.steps(400000, 400000)
    # single long run: compare with a generous 4 % band on <r3> and <U>
    assert roll[-1, 3] == pytest.approx(cf[2], rel=0.06)
    assert roll[-1, 15] == pytest.approx(cf[14], rel=0.06)
