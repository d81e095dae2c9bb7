"""Developer tuning sweep over k_run_cta variants (PMC_RUN_CFG); run on the GPU box."""
# NOTE: the PMC_*_CFG launch-shape variants exist only in tuning builds: `make -C polymer-stats_b200/csrc clean all TUNING=1`.

import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys, time
sys.path.insert(0, os.path.join(%r, "polymer-stats_b200"))
import polymc as pm
n, R, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
c = pm.make_case(n=n, E0=1.0, Fz=0.5, energy_type="interacting")
ens = pm.Ensemble(c, replicas=R, seed=20260101)
ens.run(max(10, steps // 5), 0, fetch_rows=False)
best = 1e30
for _ in range(3):
    ens.run(steps, steps, fetch_rows=False)
    best = min(best, ens.last_run_ms())
F = 2*34*((n-1)*(n-2)/6+(n-1))
print("%%d %%d %%.3f %%.4f %%.2f" %% (n, R, best, R*steps/best/1e3, R*steps/best/1e3*1e6*F/1e12))
''' % ROOT
for n, R, steps, cfgs in ((512, 4096, 300, [0, "win1", "win1286", "win2563"]),
                          (100, 8192, 2000, [0, "win1", "win648", "win1286", "thr64", "thr64win"]),
                          (256, 4096, 1000, [0, "win1", "win1286"]),
                          (64, 8192, 4000, [0, "win1"]),
                          (1024, 1184, 100, [0, "win1"])):
    for cfg in cfgs:
        env = dict(os.environ)
        if isinstance(cfg, str) and cfg.startswith("win"): env["PMC_RUN_WIN"] = cfg[3:]
        elif cfg == "thr64": env["PMC_CTA_THREADS"] = "64"
        elif cfg == "thr64win": env["PMC_CTA_THREADS"] = "64"; env["PMC_RUN_WIN"] = "1"
        elif cfg: env["PMC_RUN_CFG"] = str(cfg)
        out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps)], env=env, capture_output=True, text=True)
        print("cfg", cfg, "->", out.stdout.strip() or out.stderr.strip()[-300:], "(n R ms Mupd/s TF)", flush=True)
