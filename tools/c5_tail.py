"""C5 (n=4096) launch-tail experiment: one-CTA-per-chain kernel vs CTA pairs with the work-ordered queue, 148 chains
(one wave) and 1184 chains, at 50 and 200 trials per launch."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "polymer-stats_b200"))
import polymc as pm
kw = dict(n=4096, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.5, energy_type="interacting")
peak = max(pm.fp64_peak_probe(0, 1 << 16)[0] for _ in range(3))
F = 2 * 34 * ((4095 * 4094) / 6 + 4095)
for R in (148, 1184, 37):
    for S in (50, 200):
        for mode, order in (("0", "1"), ("1", "1"), ("1", "0")):
            os.environ["PMC_RUN_PAIR"], os.environ["PMC_PAIR_ORDER"] = mode, order
            with pm.Ensemble(pm.make_case(**kw), replicas=R, seed=20260101) as ens:
                name = ens.kernel_name()
                ens.run(S, S, fetch_rows=False)
                ms = []
                for _ in range(4 if R < 1000 else 2):
                    ens.run(S, S, fetch_rows=False)
                    ms.append(ens.last_run_ms())
                t = sum(ms) / len(ms)
                ups = R * S / (t * 1e-3)
                print(f"R={R:5d} S={S:4d} {name:24s} order={order}: {t:9.2f} ms  {ups/1e3:8.2f} k updates/s  {ups*F/1e12/peak:.3f} of DFMA peak ({peak:.1f} TF)", flush=True)
