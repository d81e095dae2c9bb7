"""Small ensembles of the clustering driver: G teams on G different trials (k_run_cta_cluster_spec) against G warps on one
trial (PMC_CLUSTER_SPEC=0) — developer tool."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "polymer-stats_b200"))
import polymc as pm
n, R, steps, et = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
c = pm.make_case(n=n, E0=1.0, Fz=0.25, energy_type=et, kappa=0.5, clustering=True, adj_ub=0.4)
ens = pm.Ensemble(c, replicas=R, seed=20260101)
ens.begin_stage(1.0)
ens.run_ex(2000, 0, fetch_rows=False)
best = 1e30
for _ in range(3):
    ens.run_ex(steps, steps, fetch_rows=False)
    best = min(best, ens.last_run_ms())
d = ens.diagnostics()
print("n=%%d R=%%d %%s: %%.3f ms  %%.3f M updates/s  (acceptance %%.3f)" %% (n, R, ens.kernel_name(), best, R*steps/best/1e3, d[:, 4].sum() / d[:, 5].sum()))
''' % ROOT
for n, et in ((100, "interacting"), (150, "interacting"), (64, "interacting")):
    for R, steps in ((25, 4000), (50, 4000), (100, 4000), (200, 4000), (500, 2000)):
        for spec in ("0", "1"):
            env = dict(os.environ, PMC_CLUSTER_SPEC=spec)
            out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps), et], env=env, capture_output=True, text=True)
            print("spec", spec, "->", out.stdout.strip() or out.stderr.strip()[-400:], flush=True)
