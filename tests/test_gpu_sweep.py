"""GPU tests of whole-study sweeps with batched aggregation — both drivers."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("driver", ["plain", "clustering", "clustering2d"])
def test_sweep_table_equals_file_pipeline_and_is_sharding_invariant(pm, tmp_path, driver):
    import aggregate_ref as REF
    from polymc import aggregate as agg
    from polymc import mcmc, mcmc_clustering, mcmc_clustering_2d, sweep
    host = {"plain": mcmc, "clustering": mcmc_clustering, "clustering2d": mcmc_clustering_2d}[driver]
    plist = []
    for e0 in (0.0, 1.0):
        for fz in (0.0, 0.5, 1.0):
            kw = dict(E0=e0, Fz=fz, num_monomers=20, energy_type="interacting", num_steps=1500)
            if driver == "clustering":
                kw.update(bend_mod=0.5, burn_in=200, burn_schedule="[10; 1]")
            if driver == "clustering2d":   # the 2-D launchers are Ising studies (2D/run/Ising_2024-11-06.jl)
                kw.update(energy_type="Ising", burn_in=200, burn_schedule="[10; 1]")
            plist.append(host.default_pargs(**kw))
    kappa = driver == "clustering"
    dims = 2 if driver == "clustering2d" else 3
    header, rows, texts = sweep.sweep_table(plist, driver=driver, runs=3, seed=5, kappaflag=kappa)
    assert len(rows) == 18 and len(rows[0]) == (9 if kappa else 8) + {"plain": 20, "clustering": 22, "clustering2d": 15}[driver]
    outdir = tmp_path / "outs"
    agg.write_out_files(str(outdir), texts)
    rheader, rrows = REF.aggregate_mcmc(str(outdir), "*.out", "dielectric", kappaflag=kappa, runflag=True, dims=dims)
    assert header == rheader
    np.testing.assert_array_equal(np.array(rows), np.array(rrows))
    ar = np.array(rows)[:, -1]
    assert np.all((ar > 0) & (ar < 1))
    # sharding by global chain id is invisible: 3 emulated ranks reproduce the single-rank result bit for bit
    cases = [host.case_from_pargs(p) for p in plist]
    proto = None if driver == "plain" else dict(burn_in=200, schedule=[10.0, 1.0])
    if driver == "clustering2d":
        assert header[-15:] == ["r1", "r2", "lambda1", "lambda2", "r1sq", "r2sq", "rsquared", "p1", "p2", "p1sq", "p2sq",
                                "psquared", "U", "Usquared", "AR"]
        assert len(open(outdir / sorted(os.listdir(outdir))[0]).read().strip().split("\n")) == 10
    one = sweep.run_shard(cases, 3, 1500, 0, 5, 0, 0, 1, proto, bit_identical=True)
    parts = []
    for gids, lo, block in one:
        parts.append((gids, block))
    ref = sweep.assemble(18, parts)
    blocks = {}
    for rank in range(3):
        for gids, lo, block in sweep.run_shard(cases, 3, 1500, 0, 5, 0, rank, 3, proto, bit_identical=True):
            blocks.setdefault(tuple(gids), []).append((lo, block))
    parts3 = []
    for gids, lst in blocks.items():
        full = np.concatenate([b for _, b in sorted(lst, key=lambda t: t[0])])
        parts3.append((np.array(gids), full))
    got = sweep.assemble(18, parts3)
    for k in ("avg", "acc_rate", "sums", "extra_sums"):
        np.testing.assert_array_equal(ref[k], got[k])


def test_run_sweep_cli(pm, tmp_path):
    out, pooled, outdir = tmp_path / "study.csv", tmp_path / "pooled.csv", tmp_path / "outs"
    argv = [sys.executable, os.path.join(ROOT, "polymer-stats_b200", "run_sweep.py"), "--driver", "clustering",
            "--out", str(out), "--pooled-out", str(pooled), "--outdir", str(outdir), "--runs", "2", "--kappaflag",
            "--by", "Fz", "--by-outdir", str(tmp_path / "by"),
            "--grid", "E0=0.5,1", "--grid", "Fz=0,0.25", "--seed", "11", "--",
            "--energy-type", "Ising", "--bend-mod", "0.5", "-n", "30", "--num-steps", "3000", "--burn-in", "300",
            "--burn-schedule", "[10; 1]", "-v", "0"]
    r = subprocess.run(argv, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = open(out).read().strip().split("\n")
    assert lines[0].startswith("E0,K1,K2,kT,Fz,Fx,n,b,kappa,r1,r2,r3,lambda1") and lines[0].endswith("Ealign,psi,AR")
    assert len(lines) == 1 + 8 and len(lines[1].split(",")) == 9 + 22
    plines = open(pooled).read().strip().split("\n")
    assert len(plines) == 1 + 4
    names = sorted(os.listdir(outdir))
    assert len(names) == 8 and names[0] == "E0-0000500_K1-0001000_K2-0000000_kT-0001000_Fz-0000000_Fx-0000000_n-0030000_b-0001000_kappa-0000500_run-001.out"
    assert len(open(outdir / names[0]).read().strip().split("\n")) == 12
    # scripts/aggregate_by.jl: one table per E0 (the sweep over Fz), 2 Fz values x 2 runs each
    by = sorted(os.listdir(tmp_path / "by"))
    assert by == ["E0-0000500_K1-0001000_K2-0000000_kT-0001000_Fx-0000000_n-0030000_b-0001000_kappa-0000500.csv",
                  "E0-0001000_K1-0001000_K2-0000000_kT-0001000_Fx-0000000_n-0030000_b-0001000_kappa-0000500.csv"]
    assert len(open(tmp_path / "by" / by[0]).read().strip().split("\n")) == 1 + 4
