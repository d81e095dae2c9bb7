"""Developer probe: where the chain-per-warp composite-trial kernel spends its time (run on the GPU box)."""
import os, sys
sys.path.insert(0, "polymer-stats_b200")
import polymc as pm
R, steps = 500, 4000
for name, kw in (("default", {}), ("no alpha carry", dict(alpha_carry=False)), ("no flips (cluster_prob=1)", dict(cluster_prob=1.0)),
                 ("kappa=0", dict(kappa=0.0)), ("always flip (cluster_prob=0)", dict(cluster_prob=0.0)),
                 ("noninteracting", dict(energy_type="noninteracting")), ("n=400", dict(n=400)), ("n=25", dict(n=25))):
    base = dict(n=100, E0=1.0, Fz=0.25, energy_type="Ising", kappa=0.5, clustering=True, adj_ub=0.4)
    base.update(kw)
    ens = pm.Ensemble(pm.make_case(**base), replicas=R, seed=1)
    ens.begin_stage(1.0)
    ens.run_ex(500, 0, fetch_rows=False)
    best = 1e30
    for _ in range(3):
        ens.run_ex(steps, steps, fetch_rows=False)
        best = min(best, ens.last_run_ms())
    cs = ens.cluster_stats()
    ar = ens.averages()[1].mean()
    print("%-30s %.3f ms  %.1f M updates/s  %.2f us per 32 trials; flips %.2f of trials, mean cluster %.2f, acceptance %.2f" % (
        name, best, R * steps / best / 1e3, best * 1e3 / (steps / 32), cs[:, 0].sum() / (R * (steps * 3 + 500)),
        cs[:, 1].sum() / max(1.0, cs[:, 0].sum()), ar))
    ens.close()
