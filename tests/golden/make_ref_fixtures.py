#!/usr/bin/env python
"""Generate the reference-pinned fixtures under tests/golden/ref/ by EXECUTING THE UNMODIFIED REFERENCE SOURCES.

    python tests/golden/make_ref_fixtures.py [--reference /root/reference]

There is no Julia in this image, so the Julia files are run by tools/minijl, a small interpreter for the Julia subset
the reference's MCMC path uses.  What runs is the reference's own code, read from /root/reference where it lies
(nothing is copied into this repository):
  * tests/golden/ref/fine_seams.jl  includes inc/eap_chain.jl, inc/average.jl, inc/acceptance.jl and records U,
    U_interaction, U_Ising, UCutoff, move!, cluster_flip! (+α), Metropolis, metropolis_acc on fixed chains
    → tests/golden/ref/fine_seams.json
  * tests/golden/ref/driver.jl  includes a whole driver script — mcmc_eap_chain.jl, mcmc_clustering_eap_chain.jl,
    2D/mcmc_clustering_eap_chain.jl — with ARGS set and `rand` scripted to pop a tape of uniforms
    → tests/golden/ref/driver_runs.json (stdout lines + the two CSV files of every run)
The same two .jl files are valid Julia: with a Julia installation, `julia tests/golden/ref/fine_seams.jl <ref> <tape>
<out.json>` / `julia tests/golden/ref/driver.jl <tape> <driver.jl> <options…>` regenerate the same fixtures from the
real runtime (libm differences: last-bit).

tests/test_reference_pin.py then checks the CPU oracle (in tape mode) against these files, and
tests/test_gpu_reference_pin.py the CUDA library against fine_seams.json.
"""
import argparse
import io
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from minijl.interp import Interp  # noqa: E402
from ref_tape import splitmix_tape, write_tape  # noqa: E402

PLAIN, CLUSTER, PLANAR = "mcmc_eap_chain.jl", "mcmc_clustering_eap_chain.jl", "2D/mcmc_clustering_eap_chain.jl"

# (name, driver script, tape seed, options)   — `--prefix` and `-v 0` are appended by the runner
DRIVER_CASES = [
    ("plain_noninteracting", PLAIN, 101, "-n 20 --energy-type noninteracting --E0 1.5 --K1 1.0 --K2 0.2 --Fz 1.0 --Fx 0.3 "
                                         "--num-steps 2000 --stepout 250 --steps-per-adjust 200"),
    ("plain_interacting", PLAIN, 102, "-n 12 --energy-type interacting --E0 1.0 --K1 1.0 --K2 0.1 --Fz 0.5 --Fx 0.2 "
                                      "--num-steps 1000 --stepout 100 --steps-per-adjust 100"),
    ("plain_ising_polar_flips", PLAIN, 103, "-n 14 --chain-type polar --mu 0.3 --energy-type Ising --E0 1.0 --Fz -0.5 --do-flips "
                                            "--num-steps 1500 --stepout 250 --steps-per-adjust 150 --kT 0.8 -b 1.2"),
    ("plain_interacting_umbrella", PLAIN, 104, "-n 8 --energy-type interacting --E0 1.2 --K1 1.0 --K2 0.3 --Fz 0.4 --umbrella-sampling "
                                               "--num-steps 800 --stepout 100 --steps-per-adjust 100"),
    ("plain_ising_three_inits", PLAIN, 105, "-n 10 --energy-type Ising --E0 1.0 --K1 1.0 --K2 0.0 --Fz 0.75 --num-inits 3 "
                                            "--num-steps 400 --stepout 100 --steps-per-adjust 100"),
    ("plain_ising_three_inits_forced", PLAIN, 106, "-n 10 --energy-type Ising --E0 1.0 --K1 1.0 --K2 0.0 --Fz 0.75 --num-inits 3 "
                                                   "--force-init --num-steps 400 --stepout 100 --steps-per-adjust 100"),
    ("cluster_ising_bend", CLUSTER, 201, "-n 10 --energy-type Ising --E0 1.0 --K1 1.0 --K2 0.1 --Fz 0.5 --bend-mod 0.5 "
                                         "--num-steps 400 --burn-in 100 --stepout 100 --steps-per-adjust 50"),
    ("cluster_interacting_bend", CLUSTER, 202, "-n 8 --energy-type interacting --E0 1.0 --K1 1.0 --K2 0.0 --Fz 0.25 --bend-mod 0.5 "
                                               "--bend-angle 0.2 --cluster-prob 0.3 --num-steps 300 --burn-in 60 --stepout 100 "
                                               "--steps-per-adjust 50"),
    ("cluster_cutoff", CLUSTER, 203, "-n 10 --energy-type cutoff --cutoff-radius 2.0 --E0 1.0 --K1 1.0 --K2 0.0 --Fz 0.25 --kT 0.7 "
                                     "--bend-mod 0.25 --num-steps 300 --burn-in 50 --stepout 100 --steps-per-adjust 50"),
    ("cluster_noninteracting_x0", CLUSTER, 204, "-n 12 --energy-type noninteracting --E0 2.0 --K1 1.0 --K2 0.0 --Fz 1.5 "
                                                "--x0 [0.5,1.0] --dx0 [0.3,0.2] --burn-schedule [10;1] --num-steps 500 --burn-in 100 "
                                                "--stepout 100 --steps-per-adjust 100"),
    ("cluster_polar_umbrella", CLUSTER, 205, "-n 8 --chain-type polar --mu 0.2 --energy-type interacting --E0 1.0 --Fz 0.5 "
                                             "--umbrella-sampling --num-steps 300 --burn-in 50 --stepout 100 --steps-per-adjust 50"),
    ("planar_ising", PLANAR, 301, "-n 10 --energy-type Ising --E0 0.5 --K1 1.0 --K2 0.1 --Fz 0.5 --num-steps 400 --burn-in 100 "
                                  "--stepout 100 --steps-per-adjust 50"),
    ("planar_interacting", PLANAR, 302, "-n 8 --energy-type interacting --E0 0.5 --K1 1.0 --K2 0.0 --Fz 0.25 --Fx 0.1 --cluster-prob 0.7 "
                                        "--num-steps 300 --burn-in 50 --stepout 100 --steps-per-adjust 50"),
    ("planar_noninteracting_polar", PLANAR, 303, "-n 12 --chain-type polar --mu 0.5 --energy-type noninteracting --E0 1.5 --Fz 1.0 "
                                                 "--num-steps 400 --burn-in 50 --stepout 100 --steps-per-adjust 100"),
]
TAPE_LEN = 60000
FINE_SEED = 7


def run_driver(ref_root, name, driver, seed, options, tmp):
    tape_path = os.path.join(tmp, f"{name}.tape")
    write_tape(tape_path, splitmix_tape(seed, TAPE_LEN))
    prefix = os.path.join(tmp, name)
    argv = options.split() + ["--prefix", prefix, "-v", "0"]
    out = io.StringIO()
    it = Interp(argv=[tape_path, os.path.join(ref_root, driver)] + argv, stdout=out, stderr=io.StringIO())
    t0 = time.time()
    it.run_file(os.path.join(HERE, "ref", "driver.jl"))
    lines = out.getvalue().strip().split("\n")
    assert lines[-1].startswith("tape_used = ")
    rec = {"name": name, "driver": driver, "tape_seed": seed, "tape_len": TAPE_LEN, "options": options.split(),
           "tape_used": int(lines[-1].split("=")[1]), "stdout": lines[:-1],
           "trajectory_csv": open(prefix + "_trajectory.csv").read(), "rolling_csv": open(prefix + "_rolling.csv").read()}
    print(f"  {name}: {time.time() - t0:.1f} s, {rec['tape_used']} uniforms", file=sys.stderr)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--only", default=None, help="regenerate one driver case into a temporary file and print its path")
    args = ap.parse_args()
    out_dir = os.path.join(HERE, "ref")
    with tempfile.TemporaryDirectory() as tmp:
        if args.only is None:
            tape_path = os.path.join(tmp, "fine.tape")
            write_tape(tape_path, splitmix_tape(FINE_SEED, TAPE_LEN))
            fine_out = os.path.join(out_dir, "fine_seams.json")
            Interp(argv=[args.reference, tape_path, fine_out]).run_file(os.path.join(out_dir, "fine_seams.jl"))
            d = json.load(open(fine_out))
            d["tape_seed"], d["tape_len"] = FINE_SEED, TAPE_LEN
            d["runtime"] = "tools/minijl (no Julia in the image) executing the unmodified files under /root/reference"
            json.dump(d, open(fine_out, "w"), indent=0)
            print(f"wrote {fine_out}: {len(d['cases'])} chains", file=sys.stderr)
        runs = [run_driver(args.reference, *c, tmp) for c in DRIVER_CASES if args.only in (None, c[0])]
        doc = {"generator": "tests/golden/make_ref_fixtures.py → tests/golden/ref/driver.jl on the unmodified reference drivers",
               "runtime": "tools/minijl (no Julia in the image)", "runs": runs}
        path = os.path.join(out_dir, "driver_runs.json") if args.only is None else os.path.join(tempfile.gettempdir(), f"ref_{args.only}.json")
        json.dump(doc, open(path, "w"), indent=0)
        print(f"wrote {path}: {len(runs)} runs", file=sys.stderr)
        if args.only:
            print(path)


if __name__ == "__main__":
    main()
