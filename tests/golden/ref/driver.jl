# driver.jl — run one of the reference's driver scripts, unmodified, on a scripted rand tape.
#
#   julia tests/golden/ref/driver.jl <tape.txt> <path/to/reference/driver.jl> <driver options...>
#
# e.g.  julia tests/golden/ref/driver.jl tape.txt /root/reference/mcmc_eap_chain.jl -n 12 --energy-type interacting ...
# The driver prints its result lines to stdout and writes <prefix>_trajectory.csv / <prefix>_rolling.csv as always.
using Distributions
include(joinpath(@__DIR__, "prelude.jl"))

const TAPE_PATH = ARGS[1]
const DRIVER_PATH = ARGS[2]
const DRIVER_ARGS = ARGS[3:end]
load_tape!(TAPE_PATH)
empty!(ARGS)
append!(ARGS, DRIVER_ARGS)
include(DRIVER_PATH)
println("tape_used = ", RAND_USED[1])
