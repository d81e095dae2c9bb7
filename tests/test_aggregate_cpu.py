"""CPU tests of the batched aggregation: the tables built in memory by polymc.aggregate
equal what scripts/aggregate_mcmc.jl / reduce_tabular_data.jl (restated in oracle/aggregate_ref.py) build from
the `.out` files on disk — bit for bit, because both go through the same Julia-style float text."""
import math
import os

import numpy as np
import pytest


def _fake_results(rng, n_cases, runs, clustering):
    out = []
    for _ in range(n_cases * runs):
        avg = rng.normal(size=16) * 10.0 ** rng.integers(-6, 7, size=16)
        ex = (float(rng.uniform(0, 100)), float(rng.uniform(0, math.pi))) if clustering else None
        out.append((avg, float(rng.uniform(0, 1)), ex))
    return out


@pytest.mark.parametrize("chain_type,clustering,kappaflag,runs", [("dielectric", False, False, 1),
                                                                   ("dielectric", True, True, 3),
                                                                   ("polar", True, False, 2)])
def test_in_memory_table_equals_the_file_pipeline(pm, tmp_path, chain_type, clustering, kappaflag, runs):
    import aggregate_ref as REF
    from polymc import aggregate as agg
    from polymc import mcmc, mcmc_clustering
    host = mcmc_clustering if clustering else mcmc
    rng = np.random.default_rng(7)
    grid = [(e0, fz, n) for e0 in (0.1, 1.0, 10.0) for fz in (0.0, 0.5, 1.2345678) for n in (100, 24)]
    plist = []
    for e0, fz, n in grid:
        kw = dict(E0=e0, Fz=fz, num_monomers=n, chain_type=chain_type, mlen=0.5, mu=0.25)
        if clustering:
            kw["bend_mod"] = 0.75
        plist.append(host.default_pargs(**kw))
    res = _fake_results(rng, len(plist), runs, clustering)
    entries, texts = [], []
    for i, p in enumerate(plist):
        for r in range(runs):
            avg, ar, ex = res[i * runs + r]
            prefix = agg.prefix_of(p, chain_type, kappaflag, run=(r + 1) if runs > 1 else None)
            entries.append((prefix, agg.output_values(avg, ar, p["mlen"], p["num-monomers"], ex)))
            texts.append((prefix, agg.out_text(avg, ar, p["mlen"], p["num-monomers"], ex)))
    header, rows = agg.aggregate_table(entries, chain_type, kappaflag, runflag=runs > 1)
    outdir = tmp_path / "outs"
    agg.write_out_files(str(outdir), texts)
    assert len(os.listdir(outdir)) == len(plist) * runs
    rheader, rrows = REF.aggregate_mcmc(str(outdir), "*.out", chain_type, kappaflag, runflag=runs > 1)
    assert header == rheader and len(rows) == len(rrows)
    for a, b in zip(rows, rrows):
        assert len(a) == len(b) == len(agg.input_headers(chain_type, kappaflag)) + (22 if clustering else 20)
        # the file pipeline sees the numbers through Julia's shortest round-trip text: identical doubles
        np.testing.assert_array_equal(np.array(a), np.array(b))
    # file-name rounding of the launchers: 1.2345678 → "0001235" → 1.235
    assert any(abs(r[header.index("Fz")] - 1.235) < 1e-12 for r in rows)
    # pooled table (reduce_tabular_data.jl)
    nparams = len(agg.input_headers(chain_type, kappaflag))
    h2, pooled = agg.reduce_table(header, rows, nparams)
    _, rpooled = REF.reduce_tabular_data(rheader, rrows, chain_type, kappaflag)
    assert len(pooled) == len(plist) == len(rpooled)
    for a, b in zip(pooled, rpooled):
        np.testing.assert_allclose(a, b, rtol=1e-15, atol=0)
    # written CSV round-trips
    path = tmp_path / "agg.csv"
    agg.write_table(str(path), header, rows)
    lines = open(path).read().strip().split("\n")
    assert lines[0].split(",") == header and len(lines) == 1 + len(rows)
    back = [float(x.replace("Inf", "inf").replace("NaN", "nan")) for x in lines[1].split(",")]
    np.testing.assert_array_equal(back, rows[0])


def test_2d_table_equals_the_file_pipeline(pm, tmp_path):
    """aggregate_mcmc.jl … 2D: 15 output columns from the 10 lines of the 2-D driver."""
    import aggregate_ref as REF
    from polymc import aggregate as agg
    from polymc import mcmc_clustering_2d as host
    rng = np.random.default_rng(3)
    plist = [host.default_pargs(E0=e0, Fz=fz, num_monomers=50, energy_type="Ising") for e0 in (0.5, 2.0) for fz in (0.0, 1.0)]
    entries, texts = [], []
    for p in plist:
        avg, ar = rng.normal(size=16) * 10.0 ** rng.integers(-3, 4, size=16), float(rng.uniform(0, 1))
        prefix = agg.prefix_of(p, "dielectric")
        entries.append((prefix, agg.output_values_2d(avg, ar, p["mlen"], p["num-monomers"])))
        texts.append((prefix, agg.out_text_2d(avg, ar, p["mlen"], p["num-monomers"])))
    outdir = tmp_path / "outs"
    agg.write_out_files(str(outdir), texts)
    header, rows = agg.aggregate_table(entries, "dielectric", dims=2)
    rheader, rrows = REF.aggregate_mcmc(str(outdir), "*.out", "dielectric", dims=2)
    assert header == rheader and len(header) == 8 + 15
    np.testing.assert_array_equal(np.array(rows), np.array(rrows))
    # the 2-D host prints the same lines from its averagers
    from polymc.mcmc import Average
    avg = np.arange(1.0, 17.0)
    vas = [Average(avg[[0, 2]], 1.0), Average(avg[[3, 5]], 1.0), Average(avg[[7, 9]], 1.0), Average(avg[[10, 12]], 1.0)]
    sas = [Average(avg[6], 1.0), Average(avg[13], 1.0), Average(avg[14], 1.0), Average(avg[15], 1.0)]
    assert host.result_lines_2d(sas, vas, 0.25, 0.5, 50) == agg.out_text_2d(avg, 0.25, 0.5, 50).strip().split("\n")


@pytest.mark.parametrize("param", ["Fz", "E0", "n", "FxFz"])
@pytest.mark.parametrize("runs", [1, 12])
def test_aggregate_by_equals_the_file_pipeline(pm, tmp_path, param, runs):
    """scripts/aggregate_by.jl: one table per combination of the other parameters (the sweep over `param`), in
    memory against the script's restatement on files — including its run-number wildcard, which keeps only runs
    000-009 of a group."""
    import aggregate_ref as REF
    from polymc import aggregate as agg
    from polymc import mcmc_clustering as host
    rng = np.random.default_rng(11)
    plist = [host.default_pargs(E0=e0, Fz=fz, Fx=fx, num_monomers=n, bend_mod=0.5)
             for e0 in (0.1, 1.0) for fz in (0.0, 0.5, 2.0) for fx in (0.0, 0.25) for n in (100, 25)]
    res = _fake_results(rng, len(plist), runs, True)
    entries, texts = [], []
    for i, p in enumerate(plist):
        for r in range(runs):
            avg, ar, ex = res[i * runs + r]
            prefix = agg.prefix_of(p, "dielectric", True, run=(r + 1) if runs > 1 else None)
            entries.append((prefix, agg.output_values(avg, ar, p["mlen"], p["num-monomers"], ex)))
            texts.append((prefix, agg.out_text(avg, ar, p["mlen"], p["num-monomers"], ex)))
    outdir = tmp_path / "outs"
    agg.write_out_files(str(outdir), texts)
    mine = agg.aggregate_by(entries, param, "dielectric", kappaflag=True, runflag=runs > 1)
    ref = REF.aggregate_by(str(outdir), param, "dielectric", kappaflag=True, runflag=runs > 1)
    assert sorted(mine) == sorted(ref) and len(mine) > 1
    swept = {"Fz": 3, "E0": 2, "n": 2, "FxFz": 6}[param]
    assert len(mine) == len(plist) // swept
    for name in mine:
        (h, rows), (rh, rrows) = mine[name], ref[name]
        assert h == rh and len(rows) == len(rrows) == swept * min(runs, 9)   # runs 010.. are lost upstream
        for a, b in zip(rows, rrows):
            np.testing.assert_array_equal(np.array(a), np.array(b))
        if param != "FxFz":   # every row of a table shares all parameters but the swept one
            cols = [k for k, x in enumerate(h[:9]) if x != param]
            assert all(all(r[k] == rows[0][k] for k in cols) for r in rows)


def test_prefix_matches_the_launchers(pm):
    from polymc import aggregate as agg
    from polymc import mcmc_clustering as mc
    p = mc.default_pargs(E0=0.5, K1=1.0, K2=0.0, kT=0.1, Fz=0.25, Fx=0.0, num_monomers=24, mlen=1.0, bend_mod=0.25)
    # run/phases-kT-small-n_2023-09-09.jl:14-16
    assert agg.prefix_of(p, "dielectric", kappaflag=True) == \
        "E0-0000500_K1-0001000_K2-0000000_kT-0000100_Fz-0000250_Fx-0000000_n-0024000_b-0001000_kappa-0000250"
    # run/Ising_2025-12-17.jl:14-16
    assert agg.prefix_of(p, "dielectric", run=7).endswith("_b-0001000_run-007")
    assert agg.fmt(0.0005) == "0000001" and agg.fmt(-0.0015) == "-000002" and agg.fmt(1.2344999) == "0001234"
    assert agg.params_from_prefix(agg.prefix_of(p, "dielectric", run=7) + ".out", runflag=True) == \
        [0.5, 1.0, 0.0, 0.1, 0.25, 0.0, 24.0, 1.0]
    q = mc.default_pargs(chain_type="polar", mu=0.01)
    assert agg.prefix_of(q, "polar").startswith("E0-0000000_mu-0000010_kT-0001000")


def test_segments_group_whole_cases_into_one_handle(pm):
    from polymc.sweep import _segments
    R = 4
    mine = np.arange(6, 23)           # starts inside case 1, ends inside case 5
    segs = list(_segments(mine, R))
    assert segs == [(0, 2, 1, 1, 2), (2, 14, 2, 3, 4), (14, 17, 5, 1, 3)]
    covered = sum(e - p for p, e, *_ in segs)
    assert covered == len(mine)
    # non-consecutive cases (another bucket in between) never share a handle
    mine = np.concatenate([np.arange(0, 8), np.arange(12, 16)])
    assert list(_segments(mine, R)) == [(0, 8, 0, 2, 4), (8, 12, 3, 1, 4)]
