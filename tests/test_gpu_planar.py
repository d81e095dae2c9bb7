"""GPU parity tests of the 2-D tree (2D/mcmc_clustering_eap_chain.jl): `planar = 1` cases on
the composite-trial kernels, through the C ABI, against the oracle and the genuine-2-vector numpy golden
vectors (tests/golden/kat_2d.json)."""
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
from test_planar_oracle import sane_rows

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ET = {"U_ni": "noninteracting", "U_int": "interacting", "U_ising": "Ising"}


@pytest.fixture(scope="module")
def kat2d():
    with open(os.path.join(GOLDEN, "kat_2d.json")) as f:
        return json.load(f)["cases"]


def test_planar_golden_energies_and_composite_trials_on_gpu(pm, kat2d):
    for case in kat2d:
        n, E = case["n"], case["E"]
        scale = E["abs_pairs"] + abs(E["U_ni"]) + 1.0
        for key, et in ET.items():
            c = pm.make_case(n=n, energy_type=et, clustering=True, planar=True, **case["par"])
            with pm.Ensemble(c, replicas=1, seed=1) as ens:
                ens.set_state(0, case["phi"], np.full(n, 0.7))       # θ is ignored by planar handles
                phi, th = ens.get_state(0)
                np.testing.assert_array_equal(phi, case["phi"])
                assert np.all(th == 0.0)
                e = ens.energy_ex(0)
                assert abs(e["U"] - E[key]) <= 1e-12 * scale, (n, key)
                assert e["Omega"] == 0.0
                r, p = ens.observables(0)
                np.testing.assert_allclose([r[0], r[2]], E["r"], rtol=1e-12, atol=1e-13)
                np.testing.assert_allclose([p[0], p[2]], E["p"], rtol=1e-12, atol=1e-13)
                assert r[1] == 0.0 and p[1] == 0.0
                for t in case["trials"]:
                    d = ens.delta_segment(0, t["idx0"], t["dphi"], 0.0, t["reflect"], t["lo0"], t["hi0"])
                    assert abs(d["dU"] - t["d" + key]) <= 1e-11 * t["scale"], (n, key, t["idx0"], t["lo0"], t["hi0"])
                    assert d["dOmega"] == 0.0
                    np.testing.assert_allclose([d["dp1"], d["dp3"]], t["dp"], rtol=1e-10, atol=1e-11)
                    if t["reflect"]:
                        assert d["log_alpha"] == pytest.approx(t["log_alpha"], rel=1e-9, abs=1e-10)


@pytest.mark.parametrize("et,n,steps", [("noninteracting", 100, 6000), ("Ising", 100, 6000), ("interacting", 48, 1500)])
@pytest.mark.parametrize("ct,umb,carry", [("dielectric", False, True), ("polar", True, False)])
def test_planar_trajectory_matches_oracle(pm, O, et, n, steps, ct, umb, carry):
    """Two stages (each from a NEW random chain, 2D/...:151) on the shared Philox stream: same clusters, same
    decisions, same rows as the oracle — up to a collapse into a singular well (see sane_rows)."""
    kw = dict(n=n, E0=0.25, K1=1.0, K2=0.2, mu=0.2, Fz=0.5, Fx=0.1, chain_type=ct, energy_type=et, clustering=True,
              planar=True, alpha_carry=carry, umbrella=umb, adj_ub=0.4, steps_per_adjust=250, theta_step=3 * math.pi / 16)
    pc, oc = both_cases(pm, O, **kw)
    compared = 0
    with pm.Ensemble(pc, replicas=3, seed=31, chain_id_base=10) as ens:
        run = O.Run(oc, 31, 12, 1)
        for mult in (5.0, 1.0):
            ens.begin_stage(mult)
            run.begin_stage(mult)
            phi, th = ens.get_state(2)
            np.testing.assert_array_equal(phi, run.chain().state()[0])     # the stage's new chain
            traj, roll, state = ens.run_ex(steps, steps // 40, want_state=True)
            ot, orl, ost = run.steps_ex(steps, steps // 40, True)
            k = min(sane_rows(traj[2]), sane_rows(ot))
            compared += k
            np.testing.assert_allclose(state[2][:k], ost[:k], rtol=0, atol=1e-12)
            np.testing.assert_allclose(traj[2][:k], ot[:k], rtol=1e-9, atol=1e-7)
            np.testing.assert_allclose(roll[2][:k, :17], orl[:k, :17], rtol=1e-9, atol=1e-7)
            assert np.all(state[2][:, 1::2] == 0.0) and np.all(traj[2][:, 2] == 0.0)
            if k == len(ot):
                cs, ocs = ens.cluster_stats()[2], run.cluster_stats()
                assert (cs[0], cs[1], cs[2]) == (ocs["ncluster"], ocs["cluster_sum"], ocs["cluster_max"])
                assert ens.diagnostics()[2][4] == run.diag()["nacc_total"]
    assert compared >= (8 if et == "noninteracting" else 3)   # singular energies collapse fast in the plane


@pytest.mark.parametrize("kw", [dict(E0=0.0, Fz=1.5), dict(E0=2.0, K1=1.0, K2=0.0, Fz=0.7, Fx=0.3),
                                dict(E0=2.0, mu=1.5, Fz=-0.5, chain_type="polar")])
def test_planar_noninteracting_matches_closed_form_on_gpu(pm, kw):
    n, R = 20, 128
    cf = CF.planar_chain_averages(n, **kw)
    c = pm.make_case(n=n, energy_type="noninteracting", clustering=True, planar=True, cluster_prob=0.0, adj_ub=0.4, **kw)
    with pm.Ensemble(c, replicas=R, seed=555) as ens:
        ens.begin_stage(1.0)
        _, roll, _ = ens.run_ex(100000, 2000)
    # averages start at step 1 from a random chain and a planar stage cannot be burnt in (every stage draws a new
    # chain, 2D/...:151): drop the first batches of the cumulative averages instead
    k = np.arange(1, roll.shape[1] + 1)
    for col in (1, 3, 4, 6, 10, 15, 16):
        cum = roll[:, :, col] * k
        batches = np.diff(np.concatenate([np.zeros((R, 1)), cum], axis=1), axis=1)[:, 5:]
        per_chain = batches.mean(axis=1)
        sem = per_chain.std(ddof=1) / math.sqrt(R)
        want = cf[col - 1]
        assert abs(per_chain.mean() - want) <= 3.0 * sem + 1e-9 * max(1.0, abs(want)), (col, per_chain.mean(), want, sem)
    assert np.all(roll[:, :, 2] == 0.0) and np.all(roll[:, :, 5] == 0.0)


def test_planar_refusals(pm):
    with pytest.raises(pm.PolymcError):
        pm.Ensemble(pm.make_case(n=10, planar=True))                                   # needs clustering
    with pytest.raises(pm.PolymcError):
        pm.Ensemble(pm.make_case(n=10, planar=True, clustering=True, kappa=0.5))       # no bending in the 2-D tree
    with pytest.raises(pm.PolymcError):
        pm.Ensemble(pm.make_case(n=10, planar=True, clustering=True, energy_type="cutoff"))
    with pm.Ensemble(pm.make_case(n=10, planar=True, clustering=True), replicas=1) as ens:
        with pytest.raises(pm.PolymcError):
            ens.init_x0([0.0, 1.0], [0.1, 0.1])


def test_planar_cli_twin_end_to_end(pm, tmp_path):
    """`2D/mcmc_clustering_eap_chain.py` = `julia 2D/mcmc_clustering_eap_chain.jl`: argv as 2D/run/Ising_2024-11-06.jl
    builds it, 10 stdout lines with 2-vectors, the 6- and 13-column CSVs."""
    prefix = str(tmp_path / "E0-0001000_K1-0001000_K2-0000000_kT-0001000_Fz-0000100_Fx-0000000_n-0000030_b-0001000")
    argv = [sys.executable, os.path.join(ROOT, "polymer-stats_b200", "2D", "mcmc_clustering_eap_chain.py"),
            "--chain-type", "dielectric", "--energy-type", "Ising", "-b", "1.0", "--E0", "1.0", "--K1", "1.0", "--K2", "0.0",
            "--kT", "1.0", "--Fz", "0.1", "--Fx", "0.0", "-n", "30", "--num-steps", "4000", "--burn-in", "500", "-v", "0",
            "--prefix", prefix, "--seed", "3"]
    out = subprocess.run(argv, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    assert [ln.split("=")[0].strip() for ln in lines] == ["<r>", "<r/nb>", "<rj2>", "<r2>", "<p>", "<pj2>", "<p2>", "<U>",
                                                          "<U2>", "AR"]
    vals = [ast.literal_eval(ln.split("=")[1].strip()) for ln in lines]
    assert len(vals[0]) == 2 and len(vals[2]) == 2 and len(vals[4]) == 2 and 0 < vals[9] < 1
    assert vals[3] == pytest.approx(sum(vals[2]), rel=1e-12)
    trj = open(prefix + "_trajectory.csv").read().strip().split("\n")
    rol = open(prefix + "_rolling.csv").read().strip().split("\n")
    assert trj[0] == "step,r1,r3,p1,p3,U" and len(trj) == 1 + 8
    assert rol[0] == "step,r1,r3,r1sq,r3sq,rsq,p1,p3,p1sq,p3sq,psq,U,Usq" and len(rol) == 1 + 8
    assert float(rol[-1].split(",")[11]) == pytest.approx(vals[7], rel=1e-12)
