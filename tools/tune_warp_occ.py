"""k_run_warp launch bounds (CTAs of 4 chains per SM): tuning build, PMC_LIB_PATH=…_tune.so — developer tool."""
import os, subprocess, sys
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tune_lane.py")).read().split("for et in")[0])
for et in ("noninteracting", "Ising"):
    for R, steps in ((2048, 20000), (16384, 20000), (65536, 5000)):
        for cfg in (30, 40, 50, 60):
            env = dict(os.environ, PMC_LANE_MODE="2", PMC_LANE_CFG=str(cfg))
            out = subprocess.run([sys.executable, "-c", child, et, "100", str(R), str(steps)], env=env, capture_output=True, text=True)
            print("k_run_warp minblocks", cfg // 10, "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
