"""Where does the per-call overhead of a sweep go?  Times handle creation / destruction and the small result calls."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "polymer-stats_b200"))
import polymc as pm
for name, kw, R in (("C3-like", dict(n=512, mu=0.1, E0=1.0, Fz=1.0, chain_type="polar", energy_type="interacting"), 4096),
                    ("C4-like", dict(n=100, E0=1.0, Fz=0.5), 16384), ("C5-like", dict(n=4096, E0=1.0, Fz=0.5, energy_type="interacting"), 148)):
    c = pm.make_case(**kw)
    for rep in range(3):
        t0 = time.perf_counter(); ens = pm.Ensemble(c, replicas=R, seed=1); t1 = time.perf_counter()
        ens.run(10, 0); t2 = time.perf_counter()
        ens.averages(); ens.accumulators(); t3 = time.perf_counter()
        ens.close(); t4 = time.perf_counter()
        print(f"{name}: create {1e3*(t1-t0):.1f} ms, run(10) {1e3*(t2-t1):.1f} ms, results {1e3*(t3-t2):.1f} ms, destroy {1e3*(t4-t3):.1f} ms", flush=True)

# the Python side of a sweep call (polymc.sweep.run_sweep) on the C4 grid
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from polymc import sweep
cases = [pm.make_case(**kw) for kw in bench.c4_grid()]
for rep in range(3):
    t0 = time.perf_counter(); res = sweep.run_sweep(cases, 1, 1000, 0, 1, 0, None); t1 = time.perf_counter()
    print(f"run_sweep(C4 grid, 1000 trials): {1e3*(t1-t0):.1f} ms", flush=True)
os.environ["PMC_TRACE_CREATE"] = "1"
ens = pm.Ensemble(cases, replicas=1, seed=1); ens.close()
