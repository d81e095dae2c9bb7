import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-stats_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, polymc as pm, oracle as O
n, R, seed = 512, 1024, 20260101
kw = dict(n=n, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.5, energy_type="interacting")
pc, oc = pm.make_case(**kw), O.make_case(**kw)
ens = pm.Ensemble(pc, replicas=R, seed=seed)
prev_drift = np.zeros(R)
phi0, th0 = ens.get_state_all()
found = 0
for s in range(1, 121):
    ens.run(1, 0)
    d = ens.diagnostics()
    # drift_max is updated at the NEXT refresh; force one via energy? refresh happens at start of run -> lag of one step
    jump = d[:, 7] - prev_drift
    bad = np.where(jump > 1e-9)[0]
    for c in bad[:3]:
        # the offending step is s-1 (drift detected at the refresh that begins step s)
        st = s - 1
        idx, up, flip, ut, eps = O.draw_step(seed, int(c), 0, st, n)
        dphi = -3*np.pi/8 + 2*(3*np.pi/8)*up; dth = -3*np.pi/16 + 2*(3*np.pi/16)*ut
        och = O.Chain(oc, phi_prev[c], th_prev[c])
        do = och.delta_u(idx, dphi, dth)
        tmp = pm.Ensemble(pc, replicas=1, seed=1)
        tmp.set_state(0, phi_prev[c], th_prev[c])
        dg = tmp.delta_u(0, idx, dphi, dth)
        tmp.close()
        print(f"step {st} chain {c} jump {jump[c]:.3e} idx {idx} oracle dU {do['dU']:.12e} gpu dU {dg['dU']:.12e} abs_sum {do['abs_sum']:.3e} diff {dg['dU']-do['dU']:.3e} dOm {do['dOmega']:.3f}")
        found += 1
    prev_drift = d[:, 7].copy()
    phi_prev, th_prev = phi0, th0
    phi0, th0 = ens.get_state_all()
print("found", found, "max drift", prev_drift.max())
