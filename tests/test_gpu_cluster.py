"""GPU parity tests of the clustering driver (mcmc_clustering_eap_chain.jl):
the composite-trial kernels, called through the C ABI v2 entry points, against the CPU oracle on the
same seeded inputs and against the golden vectors of the independent numpy restatement.

Bar: energies / changed-term sums to 1e-12·Σ|pair terms|; trajectories on the shared Philox stream take
the same cluster and accept/reject decisions (identical states, identical cluster statistics);
ensemble averages within 3σ of the oracle MCMC and of closed forms.
"""
import ast
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import closed_form as CF
from conftest import both_cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ET = {"U_ni": "noninteracting", "U_int": "interacting", "U_ising": "Ising", "U_cut_bare": "cutoff",
      "U_cut_full": "cutoff"}
TOL = 1e-12


@pytest.fixture(scope="module")
def katc():
    with open(os.path.join(GOLDEN, "kat_cluster.json")) as f:
        return json.load(f)["cases"]


def test_golden_energies_and_composite_trials_on_gpu(pm, katc):
    for case in katc:
        n, E = case["n"], case["E"]
        scale = E["abs_pairs"] + abs(E["U_ni"]) + 1.0
        for key, et in ET.items():
            c = pm.make_case(n=n, energy_type=et, clustering=True, cutoff_full=(key == "U_cut_full"), **case["par"])
            with pm.Ensemble(c, replicas=1, seed=1) as ens:
                ens.set_state(0, case["phi"], case["theta"])
                e = ens.energy_ex(0)
                assert abs(e["U"] - E[key]) <= TOL * scale, (n, key)
                assert e["su"] == pytest.approx(E["su"], rel=1e-12, abs=1e-12)
                assert e["Ubend"] == pytest.approx(E["Ubend"], rel=1e-12, abs=1e-13)
                assert e["psi"] == pytest.approx(E["psi_mean"], rel=1e-12)
                assert e["cos2"] == pytest.approx(E["cos2"], rel=1e-12)
                for t in case["trials"]:
                    tol = 1e-9 if (key.startswith("U_cut") and t["cut_pairs_changed"]) else 1e-11
                    d = ens.delta_segment(0, t["idx0"], t["dphi"], t["dtheta"], t["reflect"], t["lo0"], t["hi0"])
                    assert abs(d["dU"] - t["d" + key]) <= tol * t["scale"], (n, key, t["idx0"], t["lo0"], t["hi0"])
                    if math.isfinite(t["dOmega"]) and not (case["theta"][t["idx0"]] + t["dtheta"] <= 0.0):
                        assert d["dOmega"] == pytest.approx(t["dOmega"], rel=1e-10, abs=1e-11)
                        assert d["dbend"] == pytest.approx(t["dUbend"], rel=1e-10, abs=1e-11)
                        assert d["dpsi"] == pytest.approx(t["dpsi_sum"], rel=1e-10, abs=1e-11)
                        assert d["dcos2"] == pytest.approx(t["dcos2"], rel=1e-10, abs=1e-11)
                        np.testing.assert_allclose([d["dp1"], d["dp2"], d["dp3"]], t["dp"], rtol=1e-10, atol=1e-11)
                        if t["reflect"]:
                            assert d["log_alpha"] == pytest.approx(t["log_alpha"], rel=1e-9, abs=1e-10)


@pytest.mark.parametrize("et", ["noninteracting", "Ising", "interacting", "cutoff"])
@pytest.mark.parametrize("ct,extra", [("dielectric", dict(K1=1.0, K2=0.25)), ("polar", dict(mu=0.5))])
@pytest.mark.parametrize("n", [2, 3, 33, 64, 200, 512])
def test_segment_delta_vs_oracle(pm, O, et, ct, extra, n):
    """Changed-term sums of scripted composite trials — cluster at either end, the whole chain, a single
    monomer, clamped θ — against the oracle, on the random initial state of the shared stream."""
    kw = dict(n=n, E0=2.0, Fx=0.3, Fz=0.7, b=1.5, chain_type=ct, energy_type=et, kappa=0.6, psi0=0.25,
              cutoff_radius=3.0, clustering=True, **extra)
    pc, oc = both_cases(pm, O, **kw)
    rng = np.random.default_rng(n * 13 + len(et))
    with pm.Ensemble(pc, replicas=2, seed=11, chain_id_base=5) as ens:
        phi, th = ens.get_state(1)
        ophi, oth = O.draw_init(11, 6, 0, n)
        np.testing.assert_array_equal(phi, ophi)
        np.testing.assert_array_equal(th, oth)
        och = O.Chain(oc, phi, th)
        eg, eo = ens.energy_ex(1), och.energy_ex()
        scale = max(1.0, eo["abs_pair_sum"], abs(eo["U"]))
        for k in ("U", "su", "Udd", "Ubend"):
            assert abs(eg[k] - eo[k]) <= TOL * scale, k
        for k in ("Omega", "psi", "cos2"):
            assert eg[k] == pytest.approx(eo[k], rel=1e-12, abs=1e-12)
        segs = [(0, 0, 0, 1), (n - 1, max(0, n - 3), n - 1, 1), (n // 2, n // 2, n // 2, 0), (n // 2, 0, n - 1, 1),
                (n // 2, n // 2, n // 2, 1)]
        for _ in range(4):
            lo = int(rng.integers(0, n))
            hi = int(rng.integers(lo, n))
            segs.append((int(rng.integers(lo, hi + 1)), lo, hi, 1))
        for (idx, lo, hi, refl) in segs:
            for dth in (float(rng.uniform(-0.5, 0.5)), 10.0):
                dphi = float(rng.uniform(-1, 1))
                dg = ens.delta_segment(1, idx, dphi, dth, refl, lo, hi)
                do = och.delta_segment(idx, dphi, dth, refl, lo, hi)
                sc = max(1.0, do["abs_sum"])
                for k in ("dU", "dpair", "du", "drF", "dbend", "dpsi", "dcos2", "dp1", "dp2", "dp3"):
                    assert abs(dg[k] - do[k]) <= TOL * sc, (k, idx, lo, hi, refl, dth)
                assert dg["dOmega"] == pytest.approx(do["dOmega"], rel=1e-12, abs=1e-12, nan_ok=True)


@pytest.mark.parametrize("et,n,steps", [("noninteracting", 100, 6000), ("Ising", 100, 6000), ("interacting", 64, 2000),
                                        ("interacting", 200, 600), ("cutoff", 64, 2000)])
@pytest.mark.parametrize("ct,umb,carry", [("dielectric", False, True), ("polar", True, False)])
@pytest.mark.parametrize("packing", ["by chain count", "chain per lane"])
def test_clustering_trajectory_matches_oracle(pm, O, et, n, steps, ct, umb, carry, packing, monkeypatch):
    """The hot loop of mcmc_clustering_eap_chain.jl:267-336 through a two-stage kT ladder: same clusters,
    same decisions, same rows (8 + 19 + state) as the oracle on the shared Philox stream.  Non-interacting and
    Ising chains run one chain per WARP for ensembles this small (speculative trials resolved in order) and one
    chain per lane for large ones; both packings are held to the same sequence."""
    if packing == "chain per lane":
        if et not in ("noninteracting", "Ising"):
            pytest.skip("one packing for the CTA-per-chain energies")
        monkeypatch.setenv("PMC_LANE_CLUSTER_MODE", "1")
    kw = dict(n=n, E0=1.0, K1=1.0, K2=0.2, mu=0.6, Fz=0.5, Fx=0.1, chain_type=ct, energy_type=et, kappa=0.5, psi0=0.2,
              cutoff_radius=3.0, clustering=True, alpha_carry=carry, umbrella=umb, adj_ub=0.4, steps_per_adjust=250)
    pc, oc = both_cases(pm, O, **kw)
    with pm.Ensemble(pc, replicas=3, seed=31, chain_id_base=10) as ens:
        run = O.Run(oc, 31, 12, 1)
        for mult in (20.0, 1.0):
            ens.begin_stage(mult)
            run.begin_stage(mult * 1.0)
            traj, roll, state = ens.run_ex(steps, steps // 4, want_state=True)
            ot, orl, ost = run.steps_ex(steps, steps // 4, True)
            np.testing.assert_allclose(state[2], ost, rtol=0, atol=1e-12)
            np.testing.assert_allclose(traj[2], ot, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(roll[2], orl, rtol=1e-9, atol=1e-8)
            cs, ocs = ens.cluster_stats()[2], run.cluster_stats()
            assert (cs[0], cs[1], cs[2]) == (ocs["ncluster"], ocs["cluster_sum"], ocs["cluster_max"])
            d, od = ens.diagnostics()[2], run.diag()
            assert d[4] == od["nacc_total"] and d[5] == od["steps_total"]
            assert d[0] == pytest.approx(od["phi_step"], rel=1e-15)
            np.testing.assert_allclose(ens.extra_averages()[2], run.extra_averages(), rtol=1e-10)
        assert roll.shape == (3, 4, 19) and state.shape == (3, 4, 2 * n)
        # pmc_run on the same handle returns the first 17 rolling columns
        t17, r17 = ens.run(steps // 4, steps // 4)
        assert r17.shape == (3, 1, 17)


@pytest.mark.parametrize("case_seed", range(24))
def test_random_cases_match_oracle(pm, O, case_seed):
    """Randomised cases of the composite trial (all energies, both chain types, bending, cluster_prob 0 … 0.9,
    alpha_carry on/off, short adaptation periods so that proposal windows are cut often): the CUDA path takes the
    oracle's decisions — same final state, cluster statistics and acceptance counts."""
    rng = np.random.default_rng(1000 + case_seed)
    et = ["noninteracting", "Ising", "interacting", "cutoff"][case_seed % 4]
    sizes = [3, 9, 33, 64, 100, 170] if et in ("interacting", "cutoff") else [2, 5, 33, 64, 100, 257]
    n = sizes[(case_seed // 4) % len(sizes)]
    kw = dict(n=n, E0=float(rng.choice([0.0, 0.5, 2.0])), K1=1.0, K2=float(rng.choice([0.0, 0.3])), mu=0.5,
              Fz=float(rng.choice([0.0, 0.7])), Fx=float(rng.choice([0.0, 0.2])), kT=float(rng.choice([0.5, 1.0])),
              chain_type=str(rng.choice(["dielectric", "polar"])), energy_type=et,
              kappa=float(rng.choice([0.0, 0.5, 3.0])), psi0=float(rng.choice([0.0, 0.3])), cutoff_radius=3.0,
              clustering=True, cluster_prob=float(rng.choice([0.0, 0.3, 0.5, 0.9])),
              alpha_carry=bool(rng.integers(0, 2)), adj_ub=0.4, steps_per_adjust=int(rng.choice([37, 100, 1000])))
    steps = 1200 if et in ("interacting", "cutoff") else 4000
    pc, oc = both_cases(pm, O, **kw)
    seed = int(rng.integers(1, 10 ** 6))
    with pm.Ensemble(pc, replicas=2, seed=seed, chain_id_base=5) as ens:
        run = O.Run(oc, seed, 6, 1)
        collapsed = False
        for mult in (5.0, 1.0):
            ens.begin_stage(mult)            # the C ABI takes the kT multiplier of the ladder (:367-381) …
            run.begin_stage(mult * kw["kT"])  # … the oracle the stage's temperature
            traj, roll, state = ens.run_ex(steps, steps // 2, want_state=True)
            ot, orl, ost = run.steps_ex(steps, steps // 2, True)
            ok_rows = np.abs(ot[:, 7]) < 1e6
            if not np.all(ok_rows):
                # a singular well (no excluded volume): |U| ≳ 1e6, the acceptance test of ANY implementation has lost its
                # resolution — trajectories may part from the first such row on; the rows before it must still agree
                collapsed = True
                good = int(np.argmin(ok_rows))
                np.testing.assert_allclose(state[1][:good], ost[:good], rtol=0, atol=1e-11, err_msg=str(kw))
                np.testing.assert_allclose(traj[1][:good], ot[:good], rtol=1e-8, atol=1e-8, err_msg=str(kw))
                break
            np.testing.assert_allclose(state[1], ost, rtol=0, atol=1e-11, err_msg=str(kw))
            cs, ocs = ens.cluster_stats()[1], run.cluster_stats()
            assert (cs[0], cs[1], cs[2]) == (ocs["ncluster"], ocs["cluster_sum"], ocs["cluster_max"]), kw
            d, od = ens.diagnostics()[1], run.diag()
            assert d[4] == od["nacc_total"] and d[5] == od["steps_total"], kw
            np.testing.assert_allclose(traj[1], ot, rtol=1e-8, atol=1e-8, err_msg=str(kw))
        if collapsed:
            # Where cancellation is worst the trajectory test ends, and the changed-term sums take over on the collapsed
            # state itself (the oracle's).  Conditioning: a collapsed chain has neighbours at |r| = (b/2)|n̂_a + n̂_b| ~ 1e-3 b,
            # and a 1-ulp difference in a direction — the CUDA path reflects cluster monomers by symmetry (n̂z → −n̂z) where the
            # oracle, like the reference, evaluates cos(π − θ) — moves a 1/r³ term by 3·(b/2)/|r| ulp.  So here the bar is
            # 1e-10 of the sum of the magnitudes of the terms (1e-12 on every well-conditioned state, tests above).
            phi, th = run.chain().state()
            ens.set_state(1, phi, th)
            och = O.Chain(oc, phi, th)
            for _ in range(12):
                idx = int(rng.integers(0, n))
                lo, hi = max(0, idx - int(rng.integers(0, 4))), min(n - 1, idx + int(rng.integers(0, 4)))
                reflect = bool(rng.integers(0, 2))
                if not reflect:
                    lo = hi = idx
                dphi, dth = float(rng.uniform(-1, 1)), float(rng.uniform(-0.5, 0.5))
                dg = ens.delta_segment(1, idx, dphi, dth, reflect, lo, hi)
                do = och.delta_segment(idx, dphi, dth, int(reflect), lo, hi)
                scale = max(1.0, do["abs_sum"])
                assert abs(dg["dU"] - do["dU"]) <= 1e-10 * scale, (kw, idx, lo, hi, dg["dU"], do["dU"], scale)
                if math.isfinite(do["dOmega"]):
                    assert dg["dOmega"] == pytest.approx(do["dOmega"], rel=1e-10, abs=1e-11)


def test_matches_the_reference_sequence_of_full_recomputes(pm, O):
    """Against oracle algo 0 — deep copy, move!, refl_n! per cluster monomer with full recomputes,
    stateful acceptor with α (the reference's own sequence of operations)."""
    kw = dict(n=48, E0=1.0, K2=0.1, Fz=0.4, energy_type="interacting", kappa=0.3, clustering=True, adj_ub=0.4)
    pc, oc = both_cases(pm, O, **kw)
    with pm.Ensemble(pc, replicas=1, seed=8) as ens:
        run = O.Run(oc, 8, 0, 0)
        ens.begin_stage(1.0)
        run.begin_stage(1.0)
        traj, roll, state = ens.run_ex(1500, 300, want_state=True)
        ot, orl, ost = run.steps_ex(1500, 300, True)
        np.testing.assert_allclose(state[0], ost, rtol=0, atol=1e-12)
        np.testing.assert_allclose(traj[0], ot, rtol=1e-8, atol=1e-8)


def test_plain_driver_with_bending_uses_the_composite_kernels(pm, O):
    """kappa ≠ 0 without cluster flips (clustering=False): single-monomer trials with bending energy."""
    kw = dict(n=40, E0=1.0, Fz=0.5, energy_type="Ising", kappa=0.8, psi0=0.1, steps_per_adjust=200)
    pc, oc = both_cases(pm, O, **kw)
    with pm.Ensemble(pc, replicas=2, seed=4) as ens:
        run = O.Run(oc, 4, 1, 1)
        traj, roll = ens.run(3000, 500)
        ot, orl = run.steps(3000, 500)
        np.testing.assert_allclose(traj[1], ot, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(roll[1], orl, rtol=1e-9, atol=1e-9)
        assert ens.cluster_stats()[1, 0] == 0
        d = ens.delta_u(1, 7, 0.2, 0.1)
        och = run.chain()
        assert d["dU"] == pytest.approx(och.delta_segment(7, 0.2, 0.1, 0, 7, 7)["dU"], rel=1e-11, abs=1e-12)


def test_x0_and_stage_semantics_on_gpu(pm, O):
    kw = dict(n=32, E0=1.0, Fz=0.2, energy_type="Ising", clustering=True, phi_step=0.01, theta_step=0.005,
              steps_per_adjust=50, adj_scale=2.0, adj_ub=0.4)
    pc, oc = both_cases(pm, O, **kw)
    with pm.Ensemble(pc, replicas=2, seed=9, chain_id_base=2) as ens:
        ens.init_x0([0.0, math.pi / 2], [2 * math.pi, 0.1])
        run = O.Run(oc, 9, 3, 1)
        run.init_x0([0.0, math.pi / 2], [2 * math.pi, 0.1])
        phi, th = ens.get_state(1)
        ophi, oth = run.chain().state()
        np.testing.assert_array_equal(phi, ophi)
        np.testing.assert_array_equal(th, oth)
        assert np.all((th >= math.pi / 2) & (th < math.pi / 2 + 0.1))
        ens.begin_stage(100.0)
        ens.run_ex(200, 0)
        d = ens.diagnostics()[1]
        assert d[0] > 0.01 and d[5] == 200
        ens.begin_stage(1.0)
        d = ens.diagnostics()[1]
        assert d[0] == 0.01 and d[1] == 0.005 and d[5] == 0 and d[4] == 0      # a fresh mcmc() call
        ens.run_ex(300, 0)
        assert ens.averages()[2][1] == 300
        x0 = np.arange(64, dtype=float) * 0.01 + 0.05
        ens.init_x0(x0, [0.0, 0.0])
        phi, th = ens.get_state(0)
        np.testing.assert_array_equal(phi, x0[0::2])
        np.testing.assert_array_equal(th, x0[1::2])
        with pytest.raises(pm.PolymcError):
            ens.init_x0([0.0, 1.0, 2.0], [0.1, 0.1])
        with pytest.raises(pm.PolymcError):
            ens.delta_segment(0, 5, 0.1, 0.1, 1, 6, 9)      # idx outside [lo,hi]


def test_bending_only_chain_matches_closed_form_on_gpu(pm):
    """E0 = 0, F = 0, κ > 0, cluster-prob = 1 (no flips): ⟨Σψ/(n−1)⟩ vs the bond-angle quadrature, 3σ."""
    kappa, psi0, R = 2.0, 0.4, 256
    want = CF.bond_angle_mean(kappa, psi0)
    for et in ("noninteracting", "interacting"):   # lane kernel and CTA kernel (E0 = 0: no dipoles)
        c = pm.make_case(n=12, energy_type=et, clustering=True, cluster_prob=1.0, kappa=kappa, psi0=psi0, adj_ub=0.4)
        with pm.Ensemble(c, replicas=R, seed=2024) as ens:
            ens.begin_stage(1.0)
            ens.run_ex(5000, 0)
            ens.begin_stage(1.0)
            ens.run_ex(40000, 0)
            v = ens.extra_averages()[:, 1]
        sem = v.std(ddof=1) / math.sqrt(R)
        assert abs(v.mean() - want) <= 3.0 * sem, (et, v.mean(), want, sem)


@pytest.mark.parametrize("et,n", [("Ising", 40), ("interacting", 24)])
def test_clustering_ensemble_matches_cpu_mcmc(pm, O, et, n):
    """Statistical parity with the reference's law (alpha_carry as in acceptance.jl:30-33): GPU ensemble vs
    an independent CPU ensemble of the reference's own sequence (oracle algo 0), different seeds, 3σ."""
    kw = dict(n=n, E0=1.0, K1=1.0, K2=0.0, Fz=0.5, energy_type=et, kappa=0.4, clustering=True, adj_ub=0.4)
    pc, oc = both_cases(pm, O, **kw)
    Rg, Rc, burn, steps = 256, 40, 3000, 15000
    with pm.Ensemble(pc, replicas=Rg, seed=1001) as ens:
        ens.begin_stage(5.0)
        ens.run_ex(burn, 0)
        ens.begin_stage(1.0)
        ens.run_ex(steps, 0)
        g_avg, g_ar, _ = ens.averages()
        g_ex = ens.extra_averages()
    c_avg, c_ar, c_ex = [], [], []
    for c in range(Rc):
        run = O.Run(oc, 2002, c, 0)
        run.begin_stage(5.0)
        run.steps_ex(burn, 0)
        run.begin_stage(1.0)
        run.steps_ex(steps, 0)
        a, ar, _ = run.averages()
        c_avg.append(a)
        c_ar.append(ar)
        c_ex.append(run.extra_averages())
    c_avg, c_ar, c_ex = np.array(c_avg), np.array(c_ar), np.array(c_ex)

    def close(g, c, what):
        sg, sc = g.std(ddof=1) / math.sqrt(len(g)), c.std(ddof=1) / math.sqrt(len(c))
        assert abs(g.mean() - c.mean()) <= 3.0 * math.hypot(sg, sc) + 1e-12, (what, g.mean(), c.mean(), sg, sc)

    for k in range(14):
        close(g_avg[:, k], c_avg[:, k], k)
    close(g_ar, c_ar, "AR")
    close(g_ex[:, 0], c_ex[:, 0], "cos2")
    close(g_ex[:, 1], c_ex[:, 1], "psi")
    if et == "Ising":
        close(g_avg[:, 14], c_avg[:, 14], "U")


def test_full_size_clustering_properties(pm, O):
    """A production-sized clustering ensemble (n=400 all-pairs with bending, 1184 chains = 8 per SM):
    size-independent invariants of the running state."""
    n, R = 400, 1184
    c = pm.make_case(n=n, E0=1.0, K1=1.0, K2=0.0, Fz=0.25, energy_type="interacting", kappa=0.5, clustering=True,
                     adj_ub=0.4)
    with pm.Ensemble(c, replicas=R, seed=20260102) as ens:
        ens.begin_stage(1.0)
        traj, roll, state = ens.run_ex(200, 100, want_state=True)
        d = ens.diagnostics()
        assert np.all(d[:, 5] == 200)
        cs = ens.cluster_stats()
        assert np.all(cs[:, 0] > 50) and np.all(cs[:, 2] <= n) and np.all(cs[:, 1] >= cs[:, 0])
        phi, th = ens.get_state_all()
        np.testing.assert_array_equal(state[:, -1, 0::2], phi)       # the state row IS the final state
        np.testing.assert_array_equal(state[:, -1, 1::2], th)
        assert np.all((th >= 0) & (th <= math.pi))
        r_state = np.stack([(np.cos(phi) * np.sin(th)).sum(1), (np.sin(phi) * np.sin(th)).sum(1), np.cos(th).sum(1)], 1)
        np.testing.assert_allclose(traj[:, -1, 1:4], r_state, rtol=0, atol=1e-9)
        # running Σcos²θ, Σψ equal a recompute from the final state
        e = np.array([[ens.energy_ex(k)[q] for q in ("U", "psi", "cos2", "Udd")] for k in range(0, R, 97)])
        np.testing.assert_allclose(e[:, 2], (np.cos(th[0:R:97]) ** 2).sum(1), rtol=1e-12)
        rel = np.abs(d[0:R:97, 6] - e[:, 0]) / (1.0 + np.abs(e[:, 0]) + np.abs(e[:, 3]))
        assert np.median(rel) < 1e-11 and rel.max() < 1e-5        # conditioning of the recompute, see C2 test
        och = O.Chain(O.make_case(n=n, E0=1.0, K1=1.0, K2=0.0, Fz=0.25, energy_type="interacting", kappa=0.5,
                                  clustering=True), phi[97], th[97])
        eo = och.energy_ex()
        assert abs(e[1, 0] - eo["U"]) <= TOL * (1.0 + eo["abs_pair_sum"] + abs(eo["U"]))
        assert roll.shape == (R, 2, 19) and np.all(roll[:, :, 0] == np.array([100.0, 200.0]))
        assert np.all((roll[:, :, 17] >= 0) & (roll[:, :, 17] <= n)) and np.all((roll[:, :, 18] >= 0) & (roll[:, :, 18] <= math.pi))


def test_clustering_cli_twin_end_to_end(pm, tmp_path):
    """`mcmc_clustering_eap_chain.py` = `julia mcmc_clustering_eap_chain.jl`: argv as
    run/phases-kT-small-n_2023-09-09.jl builds it, 12 stdout lines, the two wide CSVs."""
    prefix = str(tmp_path / "E0-0001000_K1-0001000_K2-0000000_kT-0001000_Fz-0000100_Fx-0000000_n-0000030_b-0001000_kappa-0000500")
    argv = [sys.executable, os.path.join(ROOT, "polymer-stats_b200", "mcmc_clustering_eap_chain.py"),
            "--chain-type", "dielectric", "--energy-type", "interacting", "--x0", "[0.0; pi/2]", "-b", "1.0",
            "--bend-mod", "0.5", "--E0", "1.0", "--K1", "1.0", "--K2", "0.0", "--kT", "1.0", "--Fz", "0.1", "--Fx", "0.0",
            "-n", "30", "--num-steps", "4000", "--burn-in", "500", "-v", "0", "--prefix", prefix, "--stepout", "250",
            "--seed", "3"]
    out = subprocess.run(argv, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    assert len(lines) == 12
    assert [ln.split("=")[0].strip() for ln in lines] == ["<r>", "<r/nb>", "<rj2>", "<r2>", "<p>", "<pj2>", "<p2>", "<U>",
                                                          "<U2>", "<cos2(θ)>", "<ψ>", "AR"]
    vals = [ast.literal_eval(ln.split("=")[1].strip()) for ln in lines]   # aggregate_mcmc.jl:71-72
    assert 0 < vals[11] < 1 and 0 <= vals[9] <= 30 and 0 <= vals[10] <= math.pi
    trj = open(prefix + "_trajectory.csv").read().strip().split("\n")
    rol = open(prefix + "_rolling.csv").read().strip().split("\n")
    assert trj[0].startswith("step,r1,r2,r3,p1,p2,p3,U,phi1,theta1,phi2") and trj[0].endswith("mux30,muy30,muz30")
    assert len(trj) == 1 + 16 and len(trj[1].split(",")) == 8 + 5 * 30
    assert rol[0].endswith("U,Usq,Ealign,psi") and len(rol) == 1 + 16 and len(rol[1].split(",")) == 19
    assert trj[1].split(",")[0] == "250.0"
    row = [float(x) for x in trj[-1].split(",")]
    mu = np.array(row[8 + 60:]).reshape(30, 3)
    np.testing.assert_allclose(mu.sum(0), row[4:7], rtol=1e-9, atol=1e-9)   # Σμ_i of the dumped state = p
    th = np.array(row[9:8 + 60:2])
    assert float(rol[-1].split(",")[17]) == pytest.approx(vals[9], rel=1e-12)
    assert np.all((th >= 0) & (th <= math.pi))
    # refusals of the reference
    bad = subprocess.run(argv[:2] + ["--profile", "-v", "0"], capture_output=True, text=True, timeout=120)
    assert bad.returncode != 0 and "Not currently implemented" in bad.stderr


@pytest.mark.parametrize("kw", [
    dict(n=48, E0=0.6, Fz=0.3, kT=4.0, energy_type="interacting", kappa=0.5, cluster_prob=0.5),     # acceptance ≈ 0.5: many
    dict(n=100, E0=1.0, Fz=0.25, energy_type="interacting", kappa=0.5),                             # K1 / K5's chain
    dict(n=150, E0=1.2, Fz=0.1, kT=2.0, energy_type="cutoff", cutoff_radius=4.0, cluster_prob=0.3, chain_type="polar", mu=0.6),
    dict(n=40, E0=0.8, Fz=0.2, kT=3.0, energy_type="interacting", planar=True),                     # 2-D tree
])
def test_speculative_teams_are_the_sequential_chain(pm, kw):
    """k_run_cta_cluster_spec: 2, 4 or 8 one-warp teams evaluate different trials of the window at once and commit them
    in order (the first accepted trial ends a batch).  Every decision is taken on the state the sequential chain would
    show it, with the same per-team arithmetic as the one-warp kernel: acceptance counts, cluster statistics, the state
    rows and the final chains are IDENTICAL to k_run_cta_cluster<32,…>, whatever the number of teams — also at
    acceptance rates where most batches are cut short."""
    case = pm.make_case(clustering=True, adj_ub=0.4, steps_per_adjust=250, **kw)
    R, steps = 12, 3000
    ref = None
    for hint, teams in ((10 ** 6, 1), (600, 2), (300, 4), (0, 8)):
        with pm.Ensemble(case, replicas=R, seed=2024, ensemble_chains=hint) as ens:
            name = ens.kernel_name()
            assert name.startswith("k_run_cta_cluster<32," if teams == 1 else f"k_run_cta_cluster_spec<{teams},"), name
            ens.begin_stage(1.0)
            traj, roll, state = ens.run_ex(steps, 500, want_state=True)
            ens.run_ex(777, 0)                                    # a second launch: windows restart mid-adaptation-period
            got = dict(acc=ens.diagnostics()[:, 4].copy(), stats=ens.cluster_stats().copy(), final=ens.get_state_all(),
                       traj=traj, roll=roll, state=state, avg=ens.averages()[0], ex=ens.extra_averages())
        if ref is None:
            ref = got
            assert 0 < ref["acc"].sum() < R * (steps + 777)
            continue
        np.testing.assert_array_equal(got["acc"], ref["acc"])
        np.testing.assert_array_equal(got["stats"], ref["stats"])
        for a, b in zip(got["final"], ref["final"]):
            np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(got["state"], ref["state"])
        # energies: the staged positions come from a block scan whose partition follows the CTA size (rounding)
        for key in ("traj", "roll", "avg", "ex"):
            np.testing.assert_allclose(got[key], ref[key], rtol=1e-9, atol=1e-9)
