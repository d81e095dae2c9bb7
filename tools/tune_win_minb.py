import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "polymer-stats_b200"))
import polymc as pm
n, R, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
c = pm.make_case(n=n, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.5, energy_type="interacting")
ens = pm.Ensemble(c, replicas=R, seed=20260101)
ens.run(100, 0, fetch_rows=False)
best = 1e30
for _ in range(3):
    ens.run(steps, steps, fetch_rows=False)
    best = min(best, ens.last_run_ms())
print("n=%%d R=%%d: %%.3f ms  %%.3f M updates/s" %% (n, R, best, R*steps/best/1e3))
''' % ROOT
for n, R, steps in ((512, 4096, 500), (640, 4096, 300), (384, 4096, 600)):
    for w in (1, 1282, 1283, 1285, 1286):
        env = dict(os.environ, PMC_RUN_WIN=str(w), PMC_RUN_PAIR="0")
        out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps)], env=env, capture_output=True, text=True)
        print("win", w, "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
