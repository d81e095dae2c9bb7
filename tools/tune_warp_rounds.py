"""k_run_warp timing on non-interacting and Ising chains (product build, warp mode forced) — developer tool."""
import os, subprocess, sys
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tune_lane.py")).read().split("for et in")[0])
for et in ("noninteracting", "Ising"):
    for R, steps in ((500, 20000), (2048, 20000), (16384, 20000)):
        env = dict(os.environ, PMC_LANE_MODE="2")
        out = subprocess.run([sys.executable, "-c", child, et, "100", str(R), str(steps)], env=env, capture_output=True, text=True)
        print(os.environ.get("TAG", ""), "warp ->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
