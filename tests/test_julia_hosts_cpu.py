"""The three Julia hosts (polymer-stats_b200/julia/*.jl) EXECUTED without a Julia runtime and without a GPU:
tools/minijl runs them, `ccall` is marshalled through ctypes (minijl/ffi.py) into tests/mock_polymc.py, a stand-in
library whose outputs are fixed formulas.  Checked here: the option tables and the `pmc_case` the host builds equal the
Python twin's (field by field, through the package's own ctypes struct — i.e. the Julia struct has the C layout), the
call sequence is the one include/polymc.h documents, and the pooling / formatting / CSV writing of the host equals
polymc.output's (the Python twin's) on the same numbers.  The GPU counterpart (tests/test_gpu_julia_hosts.py) runs the
same hosts against the real library and compares them with the Python twins byte for byte."""
import io
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "polymer-stats_b200"))

from minijl.interp import Interp, JlError  # noqa: E402
import mock_polymc as mk  # noqa: E402

JDIR = os.path.join(ROOT, "polymer-stats_b200", "julia")


def run_host(host, argv, lib=None):
    """Run a host script like `julia host args…`; returns (stdout lines, mock library)."""
    import polymc as pm
    it = Interp(argv=list(argv))
    out = io.StringIO()
    it.stdout = out
    it.genv.vars["stdout"] = out
    mock = lib or mk.MockLib(pm.PmcCase)
    it.genv.vars["ENV"]["POLYMC_LIB"] = "mock://polymc"
    it.ffi_libs["mock://polymc"] = mock
    it.run_main(os.path.join(JDIR, host))
    return out.getvalue().splitlines(), mock


def case_fields(c):
    return {name: getattr(c, name) for name, _ in type(c)._fields_}


def read(path):
    with open(path) as f:
        return f.read().splitlines()


def expected_sums(R, cols=17):
    return np.array([[mk.acc_value(c, k) for k in range(cols)] for c in range(R)])


PLAIN = ["-n", "20", "--num-steps", "1000", "-v", "0", "--seed", "5", "--replicas", "3", "--E0", "0.7", "-J", "1.5", "-K", "0.25",
         "--kT", "0.8", "--Fz", "0.4", "-G", "0.1", "--mlen", "1.1", "-u", "interacting", "--steps-per-adjust", "300"]


@pytest.mark.parametrize("extra", [[], ["--umbrella-sampling"], ["--num-inits", "2", "--force-init"], ["--do-flips", "-T", "polar", "-m", "0.3"]])
def test_plain_host_against_the_python_twin(tmp_path, extra):
    from polymc import mcmc, output
    prefix = str(tmp_path / "jl")
    argv = PLAIN + extra + ["--prefix", prefix]
    lines, mock = run_host("polymc_host.jl", argv)
    pargs = mcmc.parse_args(argv)
    # the case the host built, read back through the package's ctypes struct == the Python twin's case
    assert case_fields(mock.cases[0]) == case_fields(mcmc.case_from_pargs(pargs))
    inits = pargs["num-inits"]
    want_calls = [("pmc_create", 1, 3, 5, 0, 0)]
    for i in range(inits):
        want_calls.append(("pmc_run", 1000, 500, 2))
        if i + 1 < inits:
            want_calls.append(("pmc_reinit",))
    assert mock.calls == want_calls + [("pmc_destroy",)]
    pooled, nrm = mcmc.pool_replicas(expected_sums(3), pargs["umbrella-sampling"])
    ar = sum(mk.diag_value(c, 4) for c in range(3)) / (3 * inits * 1000)
    assert lines == output.result_lines(pooled[:16] / nrm, ar, pargs["mlen"], pargs["num-monomers"])
    # the CSV files: header + chain 1's rows of every init, formatted like polymc.output.write_rows
    traj = read(prefix + "_trajectory.csv")
    assert traj[0] == output.TRAJ_HEADER and len(traj) == 1 + 2 * inits
    buf = io.StringIO()
    output.write_rows(buf, [[mk.traj_value(0, w, k, 500 * (w + 1)) for k in range(8)] for w in range(2)])
    assert traj[1:3] == buf.getvalue().splitlines()
    roll = read(prefix + "_rolling.csv")
    assert roll[0] == output.ROLL_HEADER and len(roll) == 1 + 2 * inits


def test_plain_host_all_gpus_of_the_box(tmp_path):
    """--devices 0: pmc_multi_create / _run / _gather / _destroy, pooled from the gathered [R][24] table."""
    from polymc import mcmc, output
    argv = PLAIN + ["--devices", "0", "--prefix", str(tmp_path / "jl")]
    lines, mock = run_host("polymc_host.jl", argv)
    names = [c[0] for c in mock.calls]
    assert names == ["pmc_multi_create", "pmc_run", "pmc_multi_gather", "pmc_destroy"]
    assert mock.calls[0][1:] == (1, 3, 5, None, 0)          # devices = C_NULL (the first ndevices), ndevices 0 = all
    nrm = np.array([mk.acc_value(c, 16) for c in range(3)])
    avg = np.array([[mk.acc_value(c, k) / nrm[c] for k in range(16)] for c in range(3)])
    pooled = (avg * nrm[:, None]).sum(axis=0) / nrm.sum()
    ar = sum(mk.diag_value(c, 4) / 1000.0 * 1000.0 for c in range(3)) / (3 * 1000)
    assert lines == output.result_lines(pooled, ar, 1.1, 20)


def test_host_reports_the_library_error(tmp_path):
    """`check(rc)`: a failing entry point becomes `error("libpolymc_b200: <pmc_last_error()> (status rc)")`."""
    import polymc as pm

    class Failing(mk.MockLib):
        def pmc_create(self, *a):
            return -2

        def pmc_last_error(self):
            return b"no CUDA device available"
    with pytest.raises(JlError, match=r"libpolymc_b200: no CUDA device available \(status -2\)"):
        run_host("polymc_host.jl", PLAIN + ["--prefix", str(tmp_path / "x")], Failing(pm.PmcCase))
    with pytest.raises(JlError, match="energy-type is not understood"):
        run_host("polymc_host.jl", ["-u", "nonsense", "--prefix", str(tmp_path / "x")])
    with pytest.raises(JlError, match="acceptance criteria has not yet been implemented"):
        run_host("polymc_host.jl", ["--acc", "kawasaki", "--prefix", str(tmp_path / "x")])


def test_real_library_fails_loudly_without_a_gpu(tmp_path):
    """Against the real libpolymc_b200.so on a box without a GPU the host stops at pmc_create with the library's message."""
    import polymc as pm
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    pm.build()
    it = Interp(argv=PLAIN + ["--prefix", str(tmp_path / "x")])
    with pytest.raises(JlError, match="libpolymc_b200: .*no CPU fallback"):
        it.run_main(os.path.join(JDIR, "polymc_host.jl"))


CLUSTER = ["-n", "6", "--num-steps", "1000", "--burn-in", "200", "-v", "0", "--seed", "5", "--replicas", "2", "--E0", "0.9",
           "--bend-mod", "0.5", "--bend-angle", "0.2", "--Fz", "0.25", "--cluster-prob", "0.4", "-u", "interacting"]


@pytest.mark.parametrize("extra", [[], ["--x0", "[0.0; π/2]", "--umbrella-sampling"], ["--burn-schedule", "[4; 1]", "--no-alpha-carry",
                                                                                     "-u", "cutoff", "--cutoff-radius", "3.5", "--cutoff-full-energy"]])
def test_clustering_host_against_the_python_twin(tmp_path, extra):
    from polymc import mcmc, mcmc_clustering as mc, output
    prefix = str(tmp_path / "jl")
    argv = CLUSTER + extra + ["--prefix", prefix]
    lines, mock = run_host("polymc_clustering_host.jl", argv)
    pargs = mc.parse_args(argv)
    assert case_fields(mock.cases[0]) == case_fields(mc.case_from_pargs(pargs))
    schedule = mc.parse_julia_vector(pargs["burn-schedule"], "burn-schedule")
    want = [("pmc_create", 1, 2, 5, 0, 0)]
    if pargs["x0"] is not None:
        want.append(("pmc_init_x0", mc.parse_julia_vector(pargs["x0"], "x0"), mc.parse_julia_vector(pargs["dx0"], "dx0")))
    for m in schedule:
        want += [("pmc_begin_stage", float(m)), ("pmc_run_ex", 200, 0, 0)]
    want += [("pmc_begin_stage", 1.0), ("pmc_run_ex", 1000, 500, 2), ("pmc_destroy",)]
    assert mock.calls == want
    extra_sums = np.array([[mk.extra_value(c, k) for k in range(2)] for c in range(2)])
    pooled, nrm = mcmc.pool_replicas(expected_sums(2), pargs["umbrella-sampling"], extra_sums)
    ar = sum(mk.diag_value(c, 4) for c in range(2)) / (2 * 1000)
    assert lines == output.result_lines_clustering(pooled[:16] / nrm, pooled[17] / nrm, pooled[18] / nrm, ar, 1.0, 6)
    traj = read(prefix + "_trajectory.csv")
    assert traj[0] == output.traj_header_clustering(6) and len(traj) == 3
    # one row: 8 columns, the 2n angles, the 3n dipoles computed by the host from the angles (inc/dipole_response.jl)
    st = np.array([mk.state_value(0, 0, k) for k in range(12)])
    mu = mc.dipoles_of(pargs, st[0::2], st[1::2])
    buf = io.StringIO()
    output.write_rows(buf, [[mk.traj_value(0, 0, k, 500) for k in range(8)] + st.tolist() + np.asarray(mu).reshape(-1).tolist()])
    assert traj[1] == buf.getvalue().splitlines()[0]
    assert read(prefix + "_rolling.csv")[0] == output.ROLL_HEADER_CLUSTERING


def test_planar_host_against_the_python_twin(tmp_path):
    from polymc import mcmc, mcmc_clustering_2d as m2, output
    prefix = str(tmp_path / "jl")
    argv = ["-n", "6", "--num-steps", "1000", "--burn-in", "200", "-v", "0", "--seed", "5", "--replicas", "2", "--E0", "0.9",
            "--Fz", "0.25", "-u", "Ising", "--prefix", prefix]
    lines, mock = run_host("polymc_clustering_2d_host.jl", argv)
    pargs = m2.parse_args(argv)
    assert case_fields(mock.cases[0]) == case_fields(m2.case_from_pargs(pargs))
    assert [c[0] for c in mock.calls] == ["pmc_create"] + ["pmc_begin_stage", "pmc_run"] * 6 + ["pmc_destroy"]
    pooled, nrm = mcmc.pool_replicas(expected_sums(2), False)
    ar = sum(mk.diag_value(c, 4) for c in range(2)) / (2 * 1000)
    assert lines == output.result_lines_2d(pooled[:16] / nrm, ar, 1.0, 6)
    assert read(prefix + "_trajectory.csv")[0] == "step,r1,r3,p1,p3,U"
    assert read(prefix + "_rolling.csv")[0] == "step,r1,r3,r1sq,r3sq,rsq,p1,p3,p1sq,p3sq,psq,U,Usq"
    assert len(read(prefix + "_rolling.csv")) == 3


def test_pair_precision_option_reaches_the_library(tmp_path):
    """--pair-precision (B200-path extension): default fp64 = PMC_PAIR_FP64, fp32 = PMC_PAIR_FP32 through
    pmc_set_pair_precision (pmc_multi_set_pair_precision with --devices); anything else is refused by both hosts."""
    from polymc import mcmc
    _, mock = run_host("polymc_host.jl", PLAIN + ["--prefix", str(tmp_path / "a")])
    assert mock.precision == 0
    _, mock = run_host("polymc_host.jl", PLAIN + ["--pair-precision", "fp32", "--prefix", str(tmp_path / "b")])
    assert mock.precision == 1
    _, mock = run_host("polymc_host.jl", PLAIN + ["--pair-precision", "fp32", "--devices", "2", "--prefix", str(tmp_path / "c")])
    assert mock.precision == 1 and mock.calls[0][0] == "pmc_multi_create"
    with pytest.raises(JlError, match="pair-precision is not understood"):
        run_host("polymc_host.jl", PLAIN + ["--pair-precision", "fp16", "--prefix", str(tmp_path / "d")])
    assert mcmc.parse_args(PLAIN + ["--pair-precision", "fp32"])["pair-precision"] == "fp32"
    assert mcmc.parse_args(PLAIN)["pair-precision"] == "fp64"
    with pytest.raises(SystemExit):
        mcmc.parse_args(PLAIN + ["--pair-precision", "fp16"])
