"""Development check of the clustering-driver kernels against the oracle (run on the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "polymer-stats_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import polymc as pm  # noqa: E402

rng = np.random.default_rng(5)
worst = {}
for et in ("interacting", "cutoff", "Ising", "noninteracting"):
    for ct in ("dielectric", "polar"):
        for n in (24, 100):
            kw = dict(n=n, E0=1.3, K1=1.0, K2=0.2, mu=0.7, Fz=0.5, Fx=0.2, b=1.1, chain_type=ct, energy_type=et,
                      kappa=0.8, psi0=0.3, cutoff_radius=2.5, clustering=True, adj_ub=0.4, steps_per_adjust=200)
            pc, oc = pm.make_case(**kw), O.make_case(**kw)
            with pm.Ensemble(pc, replicas=3, seed=77) as ens:
                phi, th = ens.get_state(1)
                och = O.Chain(oc, phi, th)
                eg, eo = ens.energy_ex(1), och.energy_ex()
                scale = max(1.0, eo["abs_pair_sum"], abs(eo["U"]))
                e_err = max(abs(eg[k] - eo[k]) for k in ("U", "su", "Udd", "Omega", "Ubend", "psi", "cos2")) / scale
                d_err = 0.0
                for (idx, lo, hi, refl) in [(5, 3, 9, 1), (0, 0, 0, 1), (n - 1, n - 4, n - 1, 1), (7, 7, 7, 0),
                                            (10, 0, n - 1, 1), (12, 12, 12, 1), (n // 2, 2, n - 3, 1)]:
                    dg = ens.delta_segment(1, idx, 0.3, -0.2, refl, lo, hi)
                    do = och.delta_segment(idx, 0.3, -0.2, refl, lo, hi)
                    sc = max(1.0, do["abs_sum"])
                    for k in ("dU", "dOmega", "dpair", "du", "drF", "dbend", "dpsi", "dcos2", "dp1", "dp2", "dp3"):
                        d_err = max(d_err, abs(dg[k] - do[k]) / sc)
                # trajectories on the shared stream, two stages
                t_err = 0.0
                run = O.Run(oc, 77, 1, 1)
                for scale_kT in (10.0, 1.0):
                    ens.begin_stage(scale_kT)
                    run.begin_stage(kw.get("kT", 1.0) * scale_kT)
                    steps = 1500 if n <= 24 else 600
                    traj, roll, st = ens.run_ex(steps, steps // 3, want_state=True)
                    ot, orl, ost = run.steps_ex(steps, steps // 3, True)
                    sc = 1.0 + np.abs(ot).max()
                    t_err = max(t_err, np.abs(traj[1] - ot).max() / sc, np.abs(st[1] - ost).max())
                    r_err = np.abs(roll[1] - orl).max() / (1.0 + np.abs(orl).max())
                    t_err = max(t_err, r_err)
                cs = ens.cluster_stats()[1]
                ocs = run.cluster_stats()
                ok_cs = (cs[0] == ocs["ncluster"] and cs[1] == ocs["cluster_sum"] and cs[2] == ocs["cluster_max"])
                ar_g = ens.averages()[1][1]
                ar_o = run.averages()[1]
                print(f"{et:15s} {ct:10s} n={n:4d}  E {e_err:.2e}  dSeg {d_err:.2e}  traj {t_err:.2e}  "
                      f"cluster-stats {'ok' if ok_cs else 'MISMATCH'} {cs}  AR {ar_g:.4f}/{ar_o:.4f}", flush=True)
                worst[(et, ct, n)] = (e_err, d_err, t_err, ok_cs)
bad = [k for k, v in worst.items() if v[0] > 1e-12 or v[1] > 1e-12 or v[2] > 1e-8 or not v[3]]
print("BAD:", bad)
sys.exit(1 if bad else 0)
