"""Find the first trial where a speculative-team run leaves the one-team trajectory (developer tool)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "polymer-stats_b200"))
import numpy as np
import polymc as pm
kw = dict(n=64, E0=0.5, Fz=0.3, kT=3.0, energy_type="interacting", kappa=0.2, cluster_prob=0.4)
R, steps, so = 96, 200000, 50
out = {}
for hint in (10 ** 6, 600, 300, 0):
    with pm.Ensemble(pm.make_case(clustering=True, adj_ub=0.4, **kw), replicas=R, seed=31337, ensemble_chains=hint) as ens:
        ens.begin_stage(1.0)
        rows = []
        for _ in range(4):
            t, r, s = ens.run_ex(steps // 4, so, want_state=True)
            rows.append(s)
        out[hint] = (np.concatenate(rows, axis=1), ens.kernel_name(), ens.diagnostics()[:, 4].copy(), ens.cluster_stats().copy())
ref = out[10 ** 6]
for hint, (s, name, acc, cs) in out.items():
    if hint == 10 ** 6:
        continue
    d = np.abs(s - ref[0]).max(axis=2)          # [chain][row]
    bad = np.argwhere(d > 0)
    print(name, "chains with a different acceptance count:", int((acc != ref[2]).sum()), "cluster stats differ:", int((cs != ref[3]).any(axis=1).sum()))
    if len(bad):
        first = bad[np.argmin(bad[:, 1])]
        c, r = int(first[0]), int(first[1])
        print("   first differing state row: chain", c, "row", r, "= step", (r + 1) * so, "; monomers that differ:",
              np.flatnonzero(np.abs(s[c, r] - ref[0][c, r]) > 0)[:12] // 2, "launch boundary every", steps // 4)

# which of the two is the sequential chain?  the CPU oracle (ΔU algorithm) on the same Philox stream, chain 84
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import oracle as O
oc = O.make_case(clustering=True, adj_ub=0.4, **kw)
run = O.Run(oc, 31337, 84, 1)
run.begin_stage(kw["kT"])     # the oracle takes the absolute kT of the stage
_, _, ost = run.steps_ex(steps, so, True)     # one call: the oracle keeps no step counter between calls
for hint, (s, name, acc, cs) in out.items():
    d = np.abs(s[84] - ost).max(axis=1)
    bad = np.flatnonzero(d > 1e-9)
    print(name, "vs oracle, chain 84: first differing row", (int(bad[0]), "= step %d" % ((int(bad[0]) + 1) * so)) if len(bad) else None,
          "; oracle accepted", run.diagnostics()["nacc_total"] if hasattr(run, "diagnostics") else "?", "GPU", acc[84])
