# prelude.jl — a scripted `rand` for running the UNMODIFIED reference sources reproducibly.
#
# The reference draws from Julia's unseeded global RNG (mcmc_eap_chain.jl:277-287, inc/eap_chain.jl:62,273,291,307).
# Defining `rand` in Main BEFORE the reference files are included shadows Base.rand for the included code, so every
# draw pops the next uniform of a tape (a text file, one Float64 per line).  The CPU oracle consumes the same tape in
# the same order (oracle/polymc_oracle.h `orc_run_new_tape`), which makes the two runs comparable trial by trial.
#
# Valid Julia (>= 1.6): `julia tests/golden/ref/driver.jl ...`.  In this repository's containers there is no Julia, so
# the same files are executed by tools/minijl, a small interpreter for the Julia subset the reference uses
# (tests/golden/make_ref_fixtures.py).
const RAND_TAPE = Any[]
const RAND_USED = Any[0]

function rand()
  RAND_USED[1] += 1
  return popfirst!(RAND_TAPE)
end
rand(d::Uniform) = d.a + (d.b - d.a) * rand()
rand(d::Uniform, n::Int) = [rand(d) for i in 1:n]
rand(r::UnitRange) = r[1 + floor(Int, rand() * length(r))]
rand(::Type{Bool}) = rand() < 0.5

function load_tape!(path)
  for line in readlines(path)
    push!(RAND_TAPE, parse(Float64, line))
  end
end
