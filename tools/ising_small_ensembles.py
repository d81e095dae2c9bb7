import os, sys
sys.path.insert(0, "polymer-stats_b200")
import polymc as pm
# PMC_LANE_CLUSTER_MODE=1 forces one chain per lane, 2 one chain per warp (default: by chain count)
for et in ("Ising", "noninteracting"):
    for cl in (1, 0):
        for R in [int(x) for x in os.environ.get("PMC_PROBE_R", "500,2048,8192").split(",")]:
            kw = dict(n=100, E0=1.0, Fz=0.25, energy_type=et)
            if cl: kw.update(kappa=0.5, clustering=True, adj_ub=0.4)
            c = pm.make_case(**kw)
            ens = pm.Ensemble(c, replicas=R, seed=1)
            steps = 4000
            if cl:
                ens.begin_stage(1.0)
                run = lambda s, so: ens.run_ex(s, so, fetch_rows=False)
            else:
                run = lambda s, so: ens.run(s, so, fetch_rows=False)
            run(500, 0)
            best = 1e30
            for _ in range(3):
                run(steps, steps)
                best = min(best, ens.last_run_ms())
            print(et, "clustering" if cl else "plain", "R=%d: %.3f ms %.2f M updates/s  (%.1f k trials/s per chain)" % (R, best, R*steps/best/1e3, steps/best))
            ens.close()
