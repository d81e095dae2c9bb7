import os, subprocess, sys
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tune_k1.py")).read().split("for n, R, steps in")[0])
for n, R, steps in ((100, 4096, 500), (64, 8192, 500), (150, 4096, 300), (25, 16384, 1000)):
    out = subprocess.run([sys.executable, "-c", child, str(n), str(R), str(steps)], capture_output=True, text=True)
    print(os.environ.get("TAG", ""), "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
