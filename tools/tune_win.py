# NOTE: the PMC_*_CFG launch-shape variants exist only in tuning builds: `make -C polymer-stats_b200/csrc clean all TUNING=1`.
import os, subprocess, sys
ROOT = "/root/repo"
child = r'''
import os, sys
sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "polymer-stats_b200"))
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
print("%d %d %.3f ms %.4f Mupd/s %.2f TF  AR %.3f" % (n, R, best, R*steps/best/1e3, R*steps/best/1e3*1e6*F/1e12, ens.averages()[1].mean()))
'''
for n, R, steps, cfgs in ((512, 4096, 300, ["1", "1285", "643", "2562"]), (256, 4096, 1000, ["1", "1285", "643"])):
    for cfg in cfgs:
        env = dict(os.environ); env["PMC_RUN_WIN"] = cfg
        out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps)], env=env, capture_output=True, text=True)
        print("win", cfg, "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
