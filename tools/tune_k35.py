import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "polymer-stats_b200"))
import polymc as pm
n, R, steps, et = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
c = pm.make_case(n=n, E0=1.0, Fz=0.25, energy_type=et, kappa=0.5, clustering=True, adj_ub=0.4, cutoff_radius=7.5)
ens = pm.Ensemble(c, replicas=R, seed=20260101)
ens.begin_stage(1.0)
ens.run_ex(100, 0, fetch_rows=False)
best = 1e30
for _ in range(3):
    ens.run_ex(steps, steps, fetch_rows=False)
    best = min(best, ens.last_run_ms())
print("%%s n=%%d R=%%d %%s: %%.3f ms  %%.3f M updates/s" %% (et, n, R, ens.kernel_name(), best, R*steps/best/1e3))
''' % ROOT
import json
SETS = json.loads(os.environ.get("TUNE_SETS", "null")) or [[100, 500, 2000, "interacting", [0, 6404, 6405, 6406, 6408]]]
for n, R, steps, et, cfgs in SETS:
    for cfg in cfgs:
        env = dict(os.environ)
        if cfg:
            env["PMC_CLUSTER_CFG"] = str(cfg)
        out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps), et], env=env, capture_output=True, text=True)
        print("cfg", cfg, "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
