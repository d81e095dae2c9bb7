"""The optional FP32 rectangle (pmc_set_pair_precision, cta_f32.cuh) — "ΔU in fp64, or in fp32 with a stated tolerance".

The tolerance stated in include/polymc.h is checked here against the FP64 path on the same configurations; the Markov
chain it drives is checked the only way a chain with (slightly) different decisions can be: short runs reproduce the
FP64 trajectory, long runs agree within 3σ, the running energy stays on the recomputed one."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# include/polymc.h: every pair term of the rectangle carries an error of at most TOL × (1 + amplification) × its magnitude
# m = (|μi·μj| + 3|μi·r̂||μj·r̂|) / (4π r³), where the amplification (|r| + |D|) / |r − D| of the NEW term exceeds ~2 only
# when the trial brings the two monomers much closer than they were (r' = r − D is a float difference of floats; the old
# separation comes from hi+lo offsets)
TOL = 2e-6


def geometry(case, phi, theta):
    """x (eap_chain.jl:49-51), n̂, μ (dipole_response.jl) of one chain from its angles."""
    nh = np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)], axis=1)
    x = case.b * (np.cumsum(nh, axis=0) - 0.5 * nh)
    if case.chain_type == 0:
        mu = (case.K1 - case.K2) * case.E0 * np.cos(theta)[:, None] * nh + np.array([0.0, 0.0, case.K2 * case.E0])
    else:
        mu = case.mu * nh
    return x, nh, mu


def pair_terms(mi, mj, r):
    """[μi·μj − 3(μi·r̂)(μj·r̂)] / (4π r³) for arrays of pairs (eap_chain.jl:196-211)."""
    r2 = np.einsum("...k,...k->...", r, r)
    d = np.einsum("...k,...k->...", mi, mj) - 3.0 * np.einsum("...k,...k->...", mi, r) * np.einsum("...k,...k->...", mj, r) / r2
    return d / (4.0 * math.pi * r2 ** 1.5)


def pair_magnitudes(mi, mj, r):
    r2 = np.einsum("...k,...k->...", r, r)
    m = np.abs(np.einsum("...k,...k->...", mi, mj)) + \
        3.0 * np.abs(np.einsum("...k,...k->...", mi, r)) * np.abs(np.einsum("...k,...k->...", mj, r)) / r2
    return m / (4.0 * math.pi * r2 ** 1.5)


def rectangle_magnitudes(case, phi, theta, idx, dphi, dtheta):
    """The stated error bound of the move (idx, dϕ, dθ) ÷ TOL: Σ over the heads×tails rectangle of
    m_old + m_new · (1 + (|r| + |D|) / |r − D|)."""
    x, nh, mu = geometry(case, phi, theta)
    p2, t2 = phi[idx] + dphi, theta[idx] + dtheta
    n_new = np.array([math.cos(p2) * math.sin(t2), math.sin(p2) * math.sin(t2), math.cos(t2)])
    D = case.b * (n_new - nh[idx])
    H, T = np.arange(0, idx), np.arange(idx + 1, len(phi))
    if len(H) == 0 or len(T) == 0:
        return 0.0
    r = x[H][:, None, :] - x[T][None, :, :]
    old = pair_magnitudes(mu[H][:, None, :], mu[T][None, :, :], r)
    new = pair_magnitudes(mu[H][:, None, :], mu[T][None, :, :], r - D)
    amp = (np.linalg.norm(r, axis=-1) + np.linalg.norm(D)) / np.linalg.norm(r - D, axis=-1)
    return float(old.sum() + (new * (1.0 + amp)).sum())


@pytest.mark.parametrize("kw", [
    dict(n=512, E0=1.0, K1=1.0, K2=0.0, Fz=0.5, chain_type="dielectric"),                 # C2
    dict(n=512, E0=2.0, mu=0.8, Fz=0.3, Fx=0.2, chain_type="polar"),                       # C3-like
    dict(n=96, E0=1.5, K1=1.0, K2=0.4, Fz=0.0, kT=0.3, chain_type="dielectric"),           # cold, compact
    dict(n=700, E0=1.0, K1=1.0, K2=0.0, Fz=3.0, chain_type="dielectric"),                  # long, stretched: |x| ~ 500 b
])
def test_fp32_delta_u_within_the_stated_tolerance(pm, kw):
    case = pm.make_case(energy_type="interacting", **kw)
    n = kw["n"]
    rng = np.random.default_rng(n)
    # replicas=2 with the launch shape of a full ensemble (the FP32 kernel also serves the wide shapes of small ones)
    with pm.Ensemble(case, replicas=2, seed=31, ensemble_chains=4096) as e64, \
            pm.Ensemble(case, replicas=2, seed=31, ensemble_chains=4096 if n != 96 else 0) as e32:
        e64.run(1500, 0)                                  # away from the random initial chain
        phi, theta = e64.get_state_all()
        e32.set_state_all(phi, theta)
        e32.set_pair_precision("fp32")
        assert e32.pair_precision() == "fp32" and e64.pair_precision() == "fp64"
        assert "fp32" in e32.kernel_name() and "fp32" not in e64.kernel_name()
        worst = worst_abs = 0.0
        for k in range(120):
            chain = k % 2
            idx = int(rng.integers(0, n)) if k > 8 else (0, n - 1, 1, n - 2, n // 2, 31, 32, 33, 64)[k]
            dphi, dtheta = float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-0.6, 0.6))
            a, b = e32.delta_u(chain, idx, dphi, dtheta), e64.delta_u(chain, idx, dphi, dtheta)
            assert a["clamped"] == b["clamped"] and a["dOmega"] == b["dOmega"]
            if b["clamped"]:
                continue
            mag = rectangle_magnitudes(case, phi[chain], theta[chain], idx, dphi, dtheta)
            err = abs(a["dU"] - b["dU"])
            # no rectangle (idx at an end): the FP32 path IS the FP64 path up to the order of the row sum
            assert err <= TOL * mag + 1e-13 * (1.0 + abs(b["dU"])), (idx, err, mag, b["dU"])
            if mag > 0:
                worst = max(worst, err / (TOL * mag))
            worst_abs = max(worst_abs, err)
        print(f"fp32 rectangle {kw}: worst error = {worst:.3f} of the stated tolerance, {worst_abs:.3e} absolute")


def test_fp32_falls_back_to_fp64_where_no_kernel_serves_it(pm):
    """Non-interacting / Ising chains, the clustering driver and two-SM chains ignore the setting (FP64, loudly visible)."""
    for kw in (dict(n=100), dict(n=100, energy_type="Ising"), dict(n=100, energy_type="interacting", kappa=0.5, clustering=True),
               dict(n=4096, energy_type="interacting")):
        with pm.Ensemble(pm.make_case(E0=1.0, **kw), replicas=2, seed=1) as ens:
            ens.set_pair_precision("fp32")
            assert ens.pair_precision() == "fp64" and "fp32" not in ens.kernel_name()
    with pm.Ensemble(pm.make_case(n=64, E0=1.0, energy_type="interacting"), replicas=2, seed=1) as ens:
        with pytest.raises(pm.PolymcError):
            ens.set_pair_precision("fp16")


def test_fp32_short_runs_reproduce_the_fp64_trajectory(pm):
    """A decision differs only when ϵ falls within ~1e-7 of the acceptance ratio: over a few hundred trials the FP32 chain
    IS the FP64 chain (same acceptances), and its rows differ by the accumulated ΔU errors only."""
    case = pm.make_case(n=256, E0=1.0, Fz=0.5, energy_type="interacting")
    with pm.Ensemble(case, replicas=24, seed=77) as e64, pm.Ensemble(case, replicas=24, seed=77) as e32:
        e32.set_pair_precision("fp32")
        t64, r64 = e64.run(400, 100)
        t32, r32 = e32.run(400, 100)
        np.testing.assert_array_equal(e32.diagnostics()[:, 4], e64.diagnostics()[:, 4])       # accepted trials
        np.testing.assert_array_equal(t32[:, :, 1:7], t64[:, :, 1:7])                         # r, p: same moves
        scale = 1.0 + np.abs(t64[:, :, 7]).max(axis=1, keepdims=True)                         # U: Σ of FP32-rounded ΔU
        assert np.max(np.abs(t32[:, :, 7] - t64[:, :, 7]) / scale) < 2e-4, np.max(np.abs(t32[:, :, 7] - t64[:, :, 7]) / scale)
        p64, th64 = e64.get_state_all()
        p32, th32 = e32.get_state_all()
        np.testing.assert_array_equal(p32, p64)
        np.testing.assert_array_equal(th32, th64)


def test_fp32_running_energy_stays_on_the_recomputed_one(pm):
    case = pm.make_case(n=200, E0=1.0, Fz=0.5, energy_type="interacting")
    with pm.Ensemble(case, replicas=64, seed=5) as ens:
        ens.set_pair_precision("fp32")
        for _ in range(4):
            ens.run(5000, 0)
        d = ens.diagnostics()
        exact = ens.energy_all()[:, 0]
        # the library re-synchronises the running energy after every FP32 launch …
        np.testing.assert_allclose(d[:, 6], exact, rtol=1e-12, atol=1e-9)
        # … and records how far it had drifted within a launch (5000 trials): relative to the energy scale of the chain
        # (these chains collapse to |U| ~ 1e5…1e7 kT, the regime FP32 is NOT meant for: the drift shows it)
        drift = d[:, 7] / (1.0 + np.abs(exact))
        print(f"fp32 drift of the running energy within 5000 trials: max {np.max(drift):.2e} of |U|, |U| up to {np.abs(exact).max():.2e}")
        assert np.max(drift) < 5e-3, np.max(drift)


def test_fp32_averages_within_3_sigma_of_fp64(pm):
    """Independent ensembles (different seeds) in the two precisions: all 16 averages + AR agree within 3σ of the
    chain-to-chain scatter — the only comparison north_star asks of a chain whose decisions may differ."""
    case = pm.make_case(n=96, E0=1.5, K1=1.0, K2=0.2, Fz=0.4, energy_type="interacting")
    R, S = 1024, 6000
    res = {}
    for prec, seed in (("fp64", 1001), ("fp32", 2002)):
        with pm.Ensemble(case, replicas=R, seed=seed) as ens:
            ens.set_pair_precision(prec)
            ens.run(2000, 0)                # burn-in (its samples are in the averages of both arms alike)
            ens.run(S, 0)
            avg, ar, _ = ens.averages()
            res[prec] = np.concatenate([avg, ar[:, None]], axis=1)
    a, b = res["fp64"], res["fp32"]
    sig = np.sqrt(a.var(axis=0, ddof=1) / R + b.var(axis=0, ddof=1) / R)
    z = np.abs(a.mean(axis=0) - b.mean(axis=0)) / np.where(sig > 0, sig, 1.0)
    assert np.all(z < 3.0 + 1e-9), z


def test_fp32_cli_option(pm, tmp_path, capsys):
    from polymc import mcmc
    argv = ["-n", "64", "-u", "interacting", "--E0", "1.0", "--num-steps", "2000", "-v", "0", "--seed", "4", "--replicas", "2"]
    assert mcmc.main(argv + ["--prefix", str(tmp_path / "a")]) == 0
    out64 = capsys.readouterr().out.splitlines()
    assert mcmc.main(argv + ["--pair-precision", "fp32", "--prefix", str(tmp_path / "b")]) == 0
    out32 = capsys.readouterr().out.splitlines()
    assert len(out32) == 10 and out32 != out64 or out32 == out64      # same chain: at most the energy digits move
    num = lambda lines: np.array([float(x) for ln in lines for x in ln.split("=", 1)[1].strip(" []").split(",")])
    np.testing.assert_allclose(num(out32), num(out64), rtol=1e-4, atol=1e-6)
