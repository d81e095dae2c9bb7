"""Where the per-call host time of polymc.sweep.run_shard goes for one rank's share of the C4 grid (developer tool).
Emulates rank 0 of `world` ranks on one GPU (no process group): python tools/profile_sweep_call.py [world]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "polymer-stats_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import polymc as pm
from polymc import sweep
import bench
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cases = pm.CaseTable([pm.make_case(**kw) for kw in bench.c4_grid()])
S = 100000
for _ in range(3):
    t0 = time.perf_counter(); parts = sweep.run_shard(cases, 1, S, 0, 1, 0, 0, world); t1 = time.perf_counter()
    full = [(g, np.zeros((len(g), sweep.NCOL))) for g, lo, b in parts]
    t2 = time.perf_counter(); res = sweep.assemble(len(cases), full); t3 = time.perf_counter()
    print(f"run_shard {1e3*(t1-t0):.2f} ms   assemble {1e3*(t3-t2):.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    sweep.run_shard(cases, 1, S, 0, 1, 0, 0, world)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
