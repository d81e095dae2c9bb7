"""The trial at which k_run_cta_cluster<32,10> and the speculative teams part ways (developer tool)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "polymer-stats_b200"))
import numpy as np
import polymc as pm
kw = dict(n=64, E0=0.5, Fz=0.3, kT=3.0, energy_type="interacting", kappa=0.2, cluster_prob=0.4)
R, c = 96, 84
res = {}
for hint in (10 ** 6, 0):
    ens = pm.Ensemble(pm.make_case(clustering=True, adj_ub=0.4, **kw), replicas=R, seed=31337, ensemble_chains=hint)
    ens.begin_stage(1.0)
    for _ in range(3):
        ens.run_ex(50000, 0, fetch_rows=False)
    ens.run_ex(15000, 0, fetch_rows=False)
    p0, t0 = ens.get_state_all()
    traj, roll, st = ens.run_ex(50, 1, want_state=True)
    res[hint] = (ens, p0[c].copy(), t0[c].copy(), traj[c], st[c])
a, b = res[10 ** 6], res[0]
print("states at step 165000 equal:", np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), " U:", a[3][0, 7], b[3][0, 7])
d = np.abs(a[4] - b[4]).max(axis=1)
k = int(np.flatnonzero(d > 0)[0])
print("first differing trial: step", 165001 + k, " U one-team before/after:", a[3][k - 1, 7] if k else None, a[3][k, 7], " U teams before/after:", b[3][k - 1, 7] if k else None, b[3][k, 7])
pre = a[4][k - 1] if k else np.stack([a[1], a[2]], axis=1).ravel()
for name, r in (("one-team", a), ("teams", b)):
    post = r[4][k]
    ch = np.flatnonzero(np.abs(post - pre) > 0) // 2
    print(name, "monomers changed by this trial:", sorted(set(ch.tolist())))
acc = b if np.abs(b[4][k] - pre).max() > 0 else a
post = acc[4][k]
phi0, th0, phi1, th1 = pre[0::2], pre[1::2], post[0::2], post[1::2]
chg = sorted(set((np.flatnonzero(np.abs(post - pre) > 0) // 2).tolist()))
lo, hi = chg[0], chg[-1]
# idx: the monomer whose phi changed (a reflection keeps phi)
idx = [m for m in chg if phi1[m] != phi0[m]]
print("segment", lo, hi, "idx candidates", idx)
if idx:
    i = idx[0]
    dphi = phi1[i] - phi0[i]
    dth = (np.pi - th1[i]) - th0[i]
    print("move: idx", i, "dphi", dphi, "dtheta", dth)
    for name, r in (("one-team handle", a), ("teams handle", b)):
        ens = r[0]
        # put the handle back on the pre-trial state of chain c and evaluate the composite trial through the seam
        p, t = ens.get_state_all()
        p[c], t[c] = phi0, th0
        ens.set_state_all(p, t)
        o = ens.delta_segment(c, i, float(dphi), float(dth), 1, lo, hi)
        print(name, o)
