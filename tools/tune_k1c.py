"""K1: rectangle unroll of the one-warp teams × launch bounds.  PMC_LIB_PATH = a TUNING build (optionally -DPMC_UR_SHORT=2)."""
import os, subprocess, sys
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tune_k1.py")).read().split("for n, R, steps in")[0])
for n, R, steps in ((100, 4096, 500), (100, 14208, 300), (150, 4096, 300)):
    for cfg in (3208, 3210, 3212):
        env = dict(os.environ, PMC_CLUSTER_CFG=str(cfg))
        out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps)], env=env, capture_output=True, text=True)
        print(os.environ.get("TAG", ""), "cfg", cfg, "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
