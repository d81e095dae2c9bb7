import sys, os
sys.path.insert(0, "polymer-stats_b200")
import polymc as pm
for R in (25, 50, 100, 148, 149, 200, 300, 444, 445, 500, 740, 741, 1000):
    for hint in (0,):
        c = pm.make_case(n=100, E0=1.0, Fz=0.25, energy_type="interacting", kappa=0.5, clustering=True, adj_ub=0.4)
        with pm.Ensemble(c, replicas=R, seed=20260101, ensemble_chains=hint) as ens:
            ens.begin_stage(1.0)
            ens.run_ex(2000, 0, fetch_rows=False)
            best = 1e30
            for _ in range(3):
                ens.run_ex(2000, 2000, fetch_rows=False)
                best = min(best, ens.last_run_ms())
            print("R=%d hint=%d %s: %.3f ms %.2f M updates/s" % (R, hint, ens.kernel_name(), best, R * 2000 / best / 1e3), flush=True)
