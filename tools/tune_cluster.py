"""Developer tuning sweep over k_run_cta_cluster variants (PMC_CLUSTER_CFG); run on the GPU box."""
# NOTE: the PMC_*_CFG launch-shape variants exist only in tuning builds: `make -C polymer-stats_b200/csrc clean all TUNING=1`.

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
ens.run_ex(max(10, steps // 5), 0, fetch_rows=False)
best = 1e30
for _ in range(3):
    ens.run_ex(steps, steps, fetch_rows=False)
    best = min(best, ens.last_run_ms())
print("%%s n=%%d R=%%d: %%.3f ms  %%.3f M updates/s" %% (et, n, R, best, R*steps/best/1e3))
''' % ROOT
for n, R, steps, et, cfgs in ((100, 4096, 500, "interacting", [0, 3216, 6408]),
                              (128, 4096, 400, "interacting", [0, 6406]),
                              (160, 4096, 400, "interacting", [0, 3212]),
                              (200, 4096, 300, "interacting", [0, 3212, 12804]),
                              (400, 2368, 200, "cutoff", [0]),
                              (100, 65536, 2000, "Ising", [0, 6403, 6404, 6408, 3208, 3212, 3216]),
                              (100, 16384, 2000, "Ising", [0, 6404, 3208, 3216]),
                              (100, 65536, 2000, "noninteracting", [0, 6404, 3216])):
    for cfg in cfgs:
        env = dict(os.environ)
        if cfg:
            env["PMC_LANE_CLUSTER_CFG" if et in ("Ising", "noninteracting") else "PMC_CLUSTER_CFG"] = str(cfg)
        out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps), et], env=env,
                             capture_output=True, text=True)
        print("cfg", cfg, "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
