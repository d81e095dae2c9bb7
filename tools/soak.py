"""Developer soak test (run on the GPU box): long runs of every kernel family; checks that every average stays finite,
that the trial counters add up, and reports the drift of the running energy against a full recompute (diag[7])."""
import sys, time
sys.path.insert(0, "polymer-stats_b200")
import numpy as np
import polymc as pm
CASES = [
    ("plain interacting n=512", dict(n=512, E0=1.0, Fz=0.5, energy_type="interacting"), 1024, 40000, False),
    ("plain interacting n=100 (small ensemble)", dict(n=100, E0=1.0, Fz=0.5, energy_type="interacting"), 100, 400000, False),
    ("plain Ising n=100 (warp)", dict(n=100, E0=1.0, Fz=0.5, energy_type="Ising"), 2048, 2000000, False),
    ("plain non-interacting n=100 (lane)", dict(n=100, E0=1.0, Fz=0.5), 65536, 200000, False),
    ("clustering interacting n=100", dict(n=100, E0=1.0, Fz=0.25, energy_type="interacting", kappa=0.5, clustering=True, adj_ub=0.4), 1184, 200000, True),
    ("clustering cut-off n=400", dict(n=400, E0=1.0, Fz=0.25, energy_type="cutoff", cutoff_radius=7.5, kappa=0.5, clustering=True, adj_ub=0.4), 592, 20000, True),
    ("clustering Ising n=100 (warp)", dict(n=100, E0=1.0, Fz=0.25, energy_type="Ising", kappa=0.5, clustering=True, adj_ub=0.4), 500, 2000000, True),
    ("clustering Ising n=100 (lane)", dict(n=100, E0=1.0, Fz=0.25, energy_type="Ising", kappa=0.5, clustering=True, adj_ub=0.4), 32768, 100000, True),
    ("planar Ising n=100 (warp)", dict(n=100, E0=0.3, Fz=0.25, energy_type="Ising", clustering=True, planar=True, adj_ub=0.4), 500, 300000, True),
]
bad = 0
for name, kw, R, steps, cl in CASES:
    ens = pm.Ensemble(pm.make_case(**kw), replicas=R, seed=99)
    if cl:
        ens.begin_stage(1.0)
    run = (lambda s: ens.run_ex(s, 0, fetch_rows=False)) if cl else (lambda s: ens.run(s, 0, fetch_rows=False))
    t0 = time.time()
    run(steps // 2)
    run(steps - steps // 2)      # the second call re-synchronises: diag[7] = drift accumulated over the first
    dt = time.time() - t0
    avg, ar, nrm = ens.averages()
    d = ens.diagnostics()
    ok = bool(np.isfinite(avg).all() and np.all(nrm == steps) and np.all(d[:, 5] == steps) and np.all((ar > 0) & (ar < 1)))
    scale = np.maximum(1.0, np.abs(d[:, 6]))
    print("%-45s %s  %d chains x %d trials in %.1f s (%.1f M updates/s); acceptance %.3f; max drift |U_run - U_recomputed| "
          "= %.2e (relative %.1e)" % (name, "ok " if ok else "BAD", R, steps, dt, R * steps / dt / 1e6, ar.mean(),
                                      d[:, 7].max(), (d[:, 7] / scale).max()), flush=True)
    bad += not ok
    ens.close()
sys.exit(1 if bad else 0)
