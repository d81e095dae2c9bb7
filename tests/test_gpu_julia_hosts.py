"""The three Julia hosts (polymer-stats_b200/julia/*.jl) run against the REAL libpolymc_b200.so on the GPU and compared
with their Python twins byte for byte: same stdout block, same two CSV files.

There is no Julia runtime in the image: tools/minijl (the Julia-subset interpreter that also executes the unmodified
reference sources for the fixtures) runs the host, and its `ccall` is marshalled through ctypes (minijl/ffi.py) — struct
layout by field order with C alignment, column-major arrays, Ref out-parameters — into the same entry points a Julia
process would bind.  Both hosts drive the same seeded Markov chains, so every digit must agree."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
JDIR = os.path.join(ROOT, "polymer-stats_b200", "julia")


def run_julia(host, argv):
    from minijl.interp import Interp
    it = Interp(argv=list(argv))
    out = io.StringIO()
    it.stdout = out
    it.genv.vars["stdout"] = out
    it.run_main(os.path.join(JDIR, host))
    return out.getvalue().splitlines()


def run_python(main, argv):
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        assert main(list(argv)) == 0
    return out.getvalue().splitlines()


def read(path):
    with open(path) as f:
        return f.read()


def both(host, main, argv, tmp_path):
    jl = run_julia(host, argv + ["--prefix", str(tmp_path / "jl")])
    py = run_python(main, argv + ["--prefix", str(tmp_path / "py")])
    assert jl == py
    for suffix in ("_trajectory.csv", "_rolling.csv"):
        a, b = read(str(tmp_path / "jl") + suffix), read(str(tmp_path / "py") + suffix)
        assert a == b and a.count("\n") >= 2
    return jl


@pytest.mark.parametrize("opts", [
    ["-n", "24", "-u", "interacting", "--E0", "1.0", "--Fz", "0.5", "--num-steps", "3000", "--replicas", "3"],
    ["-n", "40", "-u", "Ising", "-T", "polar", "-m", "0.4", "--E0", "0.8", "--do-flips", "--num-steps", "2500", "--num-inits", "2",
     "--replicas", "2", "--steps-per-adjust", "500"],
    ["-n", "30", "--E0", "1.5", "--Fz", "1.0", "-G", "0.3", "--kT", "0.7", "--umbrella-sampling", "--num-steps", "4000", "--replicas", "4",
     "--stepout", "1000"],
    ["-n", "16", "-u", "interacting", "--E0", "0.5", "--num-steps", "2000", "--numeric-type", "big", "--force-init", "--num-inits", "2"],
])
def test_plain_host_equals_python_twin(pm, tmp_path, opts):
    from polymc import mcmc
    if "--numeric-type" in opts:
        # the Python twin prints the extended-precision digits; the Julia host prints Float64 of the compensated sums:
        # compare the numbers, not the text
        jl = run_julia("polymc_host.jl", opts + ["-v", "0", "--seed", "11", "--prefix", str(tmp_path / "jl")])
        py = run_python(mcmc.main, opts + ["-v", "0", "--seed", "11", "--prefix", str(tmp_path / "py")])
        num = lambda lines: np.array([float(x) for ln in lines for x in ln.split("=", 1)[1].strip(" []").split(",")])
        np.testing.assert_allclose(num(jl), num(py), rtol=1e-13, atol=1e-300)
        return
    both("polymc_host.jl", mcmc.main, opts + ["-v", "0", "--seed", "11"], tmp_path)


def test_plain_host_over_all_gpus_equals_one_gpu(pm, tmp_path):
    """--devices 0: pmc_multi_* from the Julia host.  The chains are the same (Philox streams keyed by chain id); the
    pooled values differ from the one-device host only by the order of the pooling sums."""
    opts = ["-n", "24", "-u", "interacting", "--E0", "1.0", "--Fz", "0.5", "--num-steps", "3000", "--replicas", "6", "-v", "0", "--seed", "3"]
    one = run_julia("polymc_host.jl", opts + ["--prefix", str(tmp_path / "one")])
    many = run_julia("polymc_host.jl", opts + ["--devices", "0", "--prefix", str(tmp_path / "many")])
    num = lambda lines: np.array([float(x) for ln in lines for x in ln.split("=", 1)[1].strip(" []").split(",")])
    np.testing.assert_allclose(num(many), num(one), rtol=1e-12, atol=1e-14)
    assert read(str(tmp_path / "one_trajectory.csv")) == read(str(tmp_path / "many_trajectory.csv"))


@pytest.mark.parametrize("opts", [
    ["-n", "30", "-u", "Ising", "--E0", "1.0", "--bend-mod", "0.5", "--Fz", "0.25", "--num-steps", "3000", "--burn-in", "500", "--replicas", "3"],
    ["-n", "20", "-u", "interacting", "--E0", "0.8", "--bend-mod", "0.3", "--bend-angle", "0.2", "--num-steps", "2000", "--burn-in", "300",
     "--burn-schedule", "[10; 1]", "--x0", "[0.0; π/2]", "--replicas", "2"],
    ["-n", "40", "-u", "cutoff", "--cutoff-radius", "3.0", "--E0", "1.2", "--num-steps", "2000", "--burn-in", "200", "--burn-schedule", "[2]",
     "--umbrella-sampling", "--no-alpha-carry", "--replicas", "2", "-T", "polar", "-m", "0.5"],
])
def test_clustering_host_equals_python_twin(pm, tmp_path, opts):
    from polymc import mcmc_clustering
    both("polymc_clustering_host.jl", mcmc_clustering.main, opts + ["-v", "0", "--seed", "7"], tmp_path)


@pytest.mark.parametrize("opts", [
    ["-n", "30", "-u", "Ising", "--E0", "1.0", "--Fz", "0.25", "--num-steps", "3000", "--burn-in", "500", "--replicas", "3"],
    ["-n", "16", "-u", "interacting", "--E0", "0.6", "-T", "polar", "-m", "0.3", "--num-steps", "2000", "--burn-in", "200",
     "--burn-schedule", "[5; 1]", "--replicas", "2"],
])
def test_planar_host_equals_python_twin(pm, tmp_path, opts):
    from polymc import mcmc_clustering_2d
    both("polymc_clustering_2d_host.jl", mcmc_clustering_2d.main, opts + ["-v", "0", "--seed", "9"], tmp_path)
