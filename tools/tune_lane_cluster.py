# NOTE: the PMC_*_CFG launch-shape variants exist only in tuning builds: `make -C polymer-stats_b200/csrc clean all TUNING=1`.
import os, subprocess, sys
child = r'''
import os, sys
sys.path.insert(0, "polymer-stats_b200")
import polymc as pm
R, steps, et = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
c = pm.make_case(n=100, E0=1.0, Fz=0.25, energy_type=et, kappa=0.5, clustering=True, adj_ub=0.4)
ens = pm.Ensemble(c, replicas=R, seed=20260101)
ens.begin_stage(1.0)
ens.run_ex(300, 0, fetch_rows=False)
best = 1e30
for _ in range(3):
    ens.run_ex(steps, steps, fetch_rows=False)
    best = min(best, ens.last_run_ms())
print("%s R=%d: %.3f ms  %.1f M updates/s" % (et, R, best, R*steps/best/1e3))
'''
for et in ("Ising", "noninteracting"):
    for R in (16384, 65536, 262144):
        for cfg in (0, 6403, 6404, 6408, 3208, 3212, 3216):
            env = dict(os.environ, PMC_LANE_CLUSTER_MODE="1")
            if cfg: env["PMC_LANE_CLUSTER_CFG"] = str(cfg)
            out = subprocess.run([sys.executable, "-c", child, str(R), "2000", et], env=env, capture_output=True, text=True)
            print("cfg", cfg, "->", out.stdout.strip() or out.stderr.strip()[-200:], flush=True)
