"""Developer tuning sweep: block size against ensemble size (few chains leave SMs idle, so a chain wants more warps);
PMC_CTA_THREADS for the single-monomer kernels, PMC_CLUSTER_CFG for the composite-trial kernel.  Run on the GPU box."""
# NOTE: the PMC_*_CFG launch-shape variants exist only in tuning builds: `make -C polymer-stats_b200/csrc clean all TUNING=1`.

import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "polymer-stats_b200"))
import polymc as pm
n, R, steps, clustering = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
if clustering:
    c = pm.make_case(n=n, E0=1.0, Fz=0.25, energy_type="interacting", kappa=0.5, clustering=True, adj_ub=0.4)
else:
    c = pm.make_case(n=n, E0=1.0, Fz=0.5, energy_type="interacting")
ens = pm.Ensemble(c, replicas=R, seed=20260101)
if clustering:
    ens.begin_stage(1.0)
run = (lambda s, so: ens.run_ex(s, so, fetch_rows=False)) if clustering else (lambda s, so: ens.run(s, so, fetch_rows=False))
run(max(10, steps // 5), 0)
best = 1e30
for _ in range(3):
    run(steps, steps)
    best = min(best, ens.last_run_ms())
print("n=%%d R=%%d: %%.3f ms  %%.3f M updates/s" %% (n, R, best, R*steps/best/1e3))
''' % ROOT
plain = [(512, R, 200, 0, [0, 128, 256, 512]) for R in (64, 148, 296, 592, 1184)] + \
        [(100, R, 1000, 0, [0, 64, 128, 256]) for R in (148, 592, 2368)] + \
        [(200, R, 600, 0, [0, 128, 256, 512]) for R in (148, 592)]
clus = [(100, R, 500, 1, [0, 3212, 6406, 12804, 25602]) for R in (148, 592, 1184, 2368)] + \
       [(200, R, 300, 1, [0, 6406, 12804, 25602]) for R in (148, 592)] + \
       [(50, R, 800, 1, [0, 3212, 6406, 12804]) for R in (148, 1184)]
for n, R, steps, cl, cfgs in plain + clus:
    for cfg in cfgs:
        env = dict(os.environ)
        if cfg:
            env["PMC_CLUSTER_CFG" if cl else "PMC_CTA_THREADS"] = str(cfg)
        out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps), str(cl)], env=env,
                             capture_output=True, text=True)
        print("clustering" if cl else "plain", "cfg", cfg, "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
