"""Tiny workload for compute-sanitizer (racecheck / memcheck), one tool per gpurun call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-stats_b200"))
import polymc as pm
for et, n in (("interacting", 70), ("interacting", 200), ("Ising", 30), ("noninteracting", 30)):
    c = pm.make_case(n=n, E0=1.0, Fz=0.5, energy_type=et, steps_per_adjust=10, do_flips=True)
    with pm.Ensemble(c, replicas=3, seed=2) as ens:
        ens.run(40, 10)
        ens.delta_u(1, n // 2, 0.1, 0.2)
        ens.energy(2)
        ens.reinit()
        ens.run(20, 10)
        print(et, n, ens.averages()[1])
