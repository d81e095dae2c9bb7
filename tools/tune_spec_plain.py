"""Small ensembles of the plain driver's interacting chains: one-warp teams on different trials (k_run_cta_win_spec) against
more warps on one trial (PMC_RUN_SPEC=0) — developer tool."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "polymer-stats_b200"))
import polymc as pm
n, R, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
c = pm.make_case(n=n, E0=1.0, Fz=0.5, energy_type="interacting")
ens = pm.Ensemble(c, replicas=R, seed=20260101)
ens.run(2000, 0, fetch_rows=False)
best = 1e30
for _ in range(3):
    ens.run(steps, steps, fetch_rows=False)
    best = min(best, ens.last_run_ms())
d = ens.diagnostics()
print("n=%%d R=%%d %%s: %%.3f ms  %%.3f M updates/s  (acceptance %%.3f)" %% (n, R, ens.kernel_name(), best, R*steps/best/1e3, d[:, 4].sum() / d[:, 5].sum()))
''' % ROOT
for n in (100, 160, 48):
    for R, steps in ((25, 8000), (100, 8000), (148, 8000), (300, 4000), (444, 4000), (600, 4000), (888, 4000), (1200, 2000)):
        for spec in ("0", "1"):
            env = dict(os.environ, PMC_RUN_SPEC=spec)
            out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps)], env=env, capture_output=True, text=True)
            print("spec", spec, "->", out.stdout.strip() or out.stderr.strip()[-400:], flush=True)
