"""Developer smoke/parity script run on the GPU box (not part of the test-suite)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "polymer-stats_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import polymc as pm
import oracle as O

def both(n, **kw):
    return pm.make_case(n=n, **kw), O.make_case(n=n, **kw)

print("devices", pm.device_count())
print("fp64 probe TF/s, ms:", pm.fp64_peak_probe(0, 1 << 16))
rng = np.random.default_rng(1)
worst = 0
for et in ("noninteracting", "interacting", "Ising"):
    for ct, extra in (("dielectric", dict(K1=1.0, K2=0.25)), ("polar", dict(mu=0.5))):
        for n in (5, 64, 200, 512):
            pc, oc = both(n, E0=2.0, Fx=0.3, Fz=0.7, b=1.5, chain_type=ct, energy_type=et, **extra)
            ens = pm.Ensemble(pc, replicas=3, seed=11)
            for c in range(3):
                phi, th = ens.get_state(c)
                och = O.Chain(oc, phi, th)
                # initial random state must match the oracle's stream
                ophi, oth = O.draw_init(11, c, 0, n)
                assert np.array_equal(phi, ophi) and np.array_equal(th, oth), "init stream mismatch"
                eg, eo = ens.energy(c), och.energy()
                scale = max(1.0, och.abs_pair_sum(), abs(eo["U"]))
                err = max(abs(eg[k] - eo[k]) for k in eg) / scale
                worst = max(worst, err)
                r, p = ens.observables(c)
                err2 = max(np.abs(r - och.r()).max(), np.abs(p - och.p()).max())
                for trial in range(6):
                    idx = [0, n - 1, n // 2, int(rng.integers(n)), int(rng.integers(n)), int(rng.integers(n))][trial]
                    dphi, dth = rng.uniform(-1, 1), rng.uniform(-0.6, 0.6)
                    if trial == 4: dth = 5.0
                    dg, do = ens.delta_u(c, idx, dphi, dth), och.delta_u(idx, dphi, dth)
                    sc = max(1.0, do["abs_sum"], abs(do["dU"]))
                    e3 = abs(dg["dU"] - do["dU"]) / sc
                    e4 = abs(dg["dOmega"] - do["dOmega"]) if np.isfinite(do["dOmega"]) else 0.0
                    worst = max(worst, e3, e4)
                    if e3 > 1e-12: print("  dU mismatch", et, ct, n, c, idx, dg, do)
            print(f"{et:15s} {ct:10s} n={n:4d} energy rel err {err:.2e} obs err {err2:.2e}")
            ens.close()
print("worst normalised error", worst)

# trajectory parity with the oracle on the same Philox stream
for et, n, steps in (("noninteracting", 100, 20000), ("Ising", 100, 20000), ("interacting", 64, 5000), ("interacting", 200, 3000)):
    for flips, umb in ((False, False), (True, True)):
        pc, oc = both(n, E0=1.0, Fz=0.5, Fx=0.2, energy_type=et, do_flips=flips, umbrella=umb, steps_per_adjust=500)
        ens = pm.Ensemble(pc, replicas=2, seed=5)
        traj, roll = ens.run(steps, 500)
        avg, ar, nrm = ens.averages()
        for c in range(2):
            r = O.Run(oc, 5, c, 1)
            ot, orl = r.steps(steps, 500)
            oavg, oar, onrm = r.averages()
            dt = np.abs(traj[c] - ot).max(); dr = np.abs(roll[c] - orl).max() / max(1, np.abs(orl).max())
            print(f"traj {et:15s} n={n} flips={flips} umb={umb} c={c}: max|traj diff| {dt:.3e} roll rel {dr:.3e} AR gpu {ar[c]:.5f} cpu {oar:.5f} diag {ens.diagnostics()[c][[0,1,7]]}")
        ens.close()

# timing of the headline config
pc, _ = both(512, E0=1.0, Fz=0.5, energy_type="interacting")
for R in (592, 4096):
    ens = pm.Ensemble(pc, replicas=R, seed=20260101)
    ens.run(100, 0)
    for steps in (200, 500):
        t = time.time(); ens.run(steps, 500, fetch_rows=False); dt = time.time() - t
        ms = ens.last_run_ms()
        ups = R * steps / (ms * 1e-3)
        print(f"R={R} steps={steps} kernel {ms:.1f} ms wall {dt*1e3:.1f} ms -> {ups/1e6:.3f} M updates/s  = {ups*2988328/1e12:.2f} TF algorithmic")
    print("AR mean", ens.averages()[1].mean(), "drift", ens.diagnostics()[:, 7].max())
    ens.close()
