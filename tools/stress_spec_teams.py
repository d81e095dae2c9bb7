"""Developer stress test (run on the GPU box), PLAIN driver variant (STRESS_PLAIN=1 in the environment selects it; default: the
composite-trial kernels): the speculative one-warp teams of the composite-trial CTA kernel
(k_run_cta_cluster_spec, 2 / 4 / 8 teams by ensemble hint) against one team per chain (k_run_cta_cluster<32,…>) on random
cases — all-pairs and cut-off energies, both chain types, bending, umbrella weights, α carry on / off, 2-D.  On the shared
Philox stream they must take the same decisions and end in identical states."""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(%r, "polymer-stats_b200"))
import polymc as pm
kw = json.loads(sys.argv[1]); R, steps, seed = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ens = pm.Ensemble(pm.make_case(**kw), replicas=R, seed=seed, ensemble_chains=int(os.environ['STRESS_HINT']))
name = ens.kernel_name()
out = []
if os.environ.get("STRESS_PLAIN") == "1":
    for _ in range(2):
        traj, roll = ens.run(steps, steps // 3)
        out.append((traj, roll, np.zeros(1)))
    cs = np.zeros(1)
else:
    for mult in (10.0, 1.0):
        ens.begin_stage(mult)
        traj, roll, state = ens.run_ex(steps, steps // 3, want_state=True)
        out.append((traj, roll, state))
    cs = ens.cluster_stats()
phi, th = ens.get_state_all()
np.savez(sys.argv[5], phi=phi, th=th, traj=out[1][0], roll=out[1][1], state=out[1][2], cs=cs,
         ar=ens.averages()[1], diag=ens.diagnostics(), name=np.array(name))
''' % ROOT
rng = np.random.default_rng(7)
bad = 0
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    planar = bool(rng.integers(0, 4) == 0)
    kw = dict(n=int(rng.choice([2, 3, 5, 17, 33, 64, 100, 131, 160])), E0=float(rng.choice([0.0, 0.5, 2.0])),
              K1=1.0, K2=float(rng.choice([0.0, 0.3])), mu=0.5, Fz=float(rng.choice([0.0, 0.7])), Fx=float(rng.choice([0.0, 0.2])),
              kT=float(rng.choice([0.3, 1.0])), chain_type=str(rng.choice(["dielectric", "polar"])),
              energy_type=str(rng.choice(["interacting", "cutoff"])), cutoff_radius=float(rng.choice([2.5, 7.5])),
              cutoff_full=bool(rng.integers(0, 2)), kappa=float(rng.choice([0.0, 0.5, 3.0])),
              psi0=float(rng.choice([0.0, 0.3])), clustering=True, cluster_prob=float(rng.choice([0.0, 0.3, 0.5, 0.9])),
              alpha_carry=bool(rng.integers(0, 2)), umbrella=bool(rng.integers(0, 4) == 0), adj_ub=0.4,
              steps_per_adjust=int(rng.choice([50, 77, 1000])))
    if os.environ.get("STRESS_PLAIN") == "1":   # mcmc_eap_chain.jl: single-monomer trials, all-pairs energy
        planar = False
        kw = dict(n=int(rng.choice([2, 3, 5, 17, 33, 48, 64, 100, 110])), E0=float(rng.choice([0.0, 0.5, 2.0])), K1=1.0,
                  K2=float(rng.choice([0.0, 0.3])), mu=0.5, Fz=float(rng.choice([0.0, 0.7])), Fx=float(rng.choice([0.0, 0.2])),
                  kT=float(rng.choice([0.3, 1.0, 5.0])), chain_type=str(rng.choice(["dielectric", "polar"])), energy_type="interacting",
                  do_flips=bool(rng.integers(0, 2)), umbrella=bool(rng.integers(0, 4) == 0),
                  steps_per_adjust=int(rng.choice([50, 77, 1000])))
    if planar:
        kw.update(planar=True, umbrella=False, kappa=0.0, psi0=0.0, energy_type="interacting")  # the 2-D tree: no bending, no cut-off
    R, steps, seed = int(rng.choice([1, 3, 7])), int(rng.choice([300, 999, 2000])), int(rng.integers(1, 10 ** 6))
    res = {}
    hints = ("1000000", str(rng.choice([800 if os.environ.get("STRESS_PLAIN") == "1" else 600, 300, 0])))
    for mode in hints:
        f = "/tmp/stress_%s.npz" % mode
        env = dict(os.environ, STRESS_HINT=mode)
        p = subprocess.run([sys.executable, "-c", child, json.dumps(kw), str(R), str(steps), str(seed), f], env=env,
                           capture_output=True, text=True)
        if p.returncode:
            print("case", t, "mode", mode, "FAILED:", p.stderr.strip()[-300:], kw)
            bad += 1
            break
        res[mode] = dict(np.load(f))
    if len(res) < 2:
        continue
    a, b = res[hints[0]], res[hints[1]]
    finite = np.isfinite(a["traj"]).all()
    ok = (np.array_equal(a["phi"], b["phi"]) and np.array_equal(a["th"], b["th"]) and np.array_equal(a["state"], b["state"])
          and np.array_equal(a["cs"], b["cs"]) and np.array_equal(a["ar"], b["ar"])
          and np.allclose(a["traj"], b["traj"], rtol=1e-9, atol=1e-9, equal_nan=True)
          and np.allclose(a["roll"], b["roll"], rtol=1e-8, atol=1e-8, equal_nan=True))
    print("case %2d %s n=%d R=%d steps=%d %s planar=%d finite=%d  %s vs %s" % (t, "ok " if ok else "MISMATCH", kw["n"], R, steps,
          kw["energy_type"], planar, finite, str(a["name"]), str(b["name"])), flush=True)
    if not ok:
        bad += 1
        print("   ", kw, seed)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
