"""Launch bound of the CTA-pair kernel against the shared-memory fit (PMC_PAIR_MINB)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "polymer-stats_b200"))
import polymc as pm
peak = max(pm.fp64_peak_probe(0, 1 << 16)[0] for _ in range(3))
for n, S in ((512, 400), (1024, 200), (1536, 100), (2048, 60)):
    kw = dict(n=n, E0=1.0, K1=1.0, K2=0.0, kT=1.0, b=1.0, Fz=0.5, energy_type="interacting")
    F = 2 * 34 * (((n - 1) * (n - 2)) / 6 + (n - 1))
    for R in (512, 2048):
        for mb in (0, 1, 2, 3, 4):
            os.environ["PMC_RUN_PAIR"] = "2"
            os.environ["PMC_PAIR_MINB"] = str(mb)
            with pm.Ensemble(pm.make_case(**kw), replicas=R, seed=20260101) as ens:
                name = ens.kernel_name()
                ens.run(S, S, fetch_rows=False)
                ms = []
                for _ in range(3):
                    ens.run(S, S, fetch_rows=False)
                    ms.append(ens.last_run_ms())
                t = min(ms)
                ups = R * S / (t * 1e-3)
                print(f"n={n:5d} R={R:5d} minb={mb} {name:26s}: {t:9.2f} ms  {ups/1e6:8.3f} M updates/s  {ups*F/1e12/peak:.3f} of DFMA peak", flush=True)
