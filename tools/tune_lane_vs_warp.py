"""Chain per lane vs chain per warp on large O(1)-ΔU ensembles (product build; PMC_LANE_MODE forces the kernel)."""
import os, subprocess, sys
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tune_lane.py")).read().split("for et in")[0])
for et in ("noninteracting", "Ising"):
    for R, steps in ((16384, 20000), (65536, 5000), (262144, 2000), (1048576, 500)):
        for mode in (1, 2):
            env = dict(os.environ, PMC_LANE_MODE=str(mode))
            out = subprocess.run([sys.executable, "-c", child, et, "100", str(R), str(steps)], env=env, capture_output=True, text=True)
            print({1: "lane", 2: "warp"}[mode], "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
