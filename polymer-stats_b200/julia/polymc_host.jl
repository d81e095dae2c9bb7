# polymc_host.jl — Julia host for libpolymc_b200.so (include/polymc.h).
#
# Drop-in for `julia mcmc_eap_chain.jl ...`: same command line (ArgParse names/defaults of
# mcmc_eap_chain.jl:19-153), same 10 stdout lines (:386-395), same two CSV files (:256-259), but the
# body of `mcmc(nsteps, pargs)` (:171-376) is a handful of `ccall`s into the CUDA library.
#
# NOTE: Julia is not installed in the image this repository is built and tested in.  This file IS executed there all the
# same: tools/minijl (the Julia-subset interpreter that also runs the unmodified reference sources for the fixtures) runs it
# with `ccall` marshalled through ctypes into the real libpolymc_b200.so — `python -m minijl polymc_host.jl <options>` with
# tools/ on PYTHONPATH.  tests/test_gpu_julia_hosts.py: stdout and both CSV files equal the Python twin's
# (../polymc/mcmc.py) byte for byte on the GPU; tests/test_julia_hosts_cpu.py: against a stand-in library without a GPU
# (case struct field by field, call sequence, pooling, formatting).  See INTEGRATION.md.
using ArgParse, Printf, DelimitedFiles, Logging

const LIBPOLYMC = get(ENV, "POLYMC_LIB", joinpath(@__DIR__, "..", "libpolymc_b200.so"))

# mirror of `pmc_case` (include/polymc.h) — field order and types must match exactly
struct PmcCase
  E0::Cdouble; K1::Cdouble; K2::Cdouble; mu::Cdouble; kT::Cdouble; Fz::Cdouble; Fx::Cdouble; b::Cdouble
  phi_step::Cdouble; theta_step::Cdouble
  adj_lb::Cdouble; adj_ub::Cdouble; adj_scale::Cdouble
  n::Int64; steps_per_adjust::Int64
  chain_type::Int32; energy_type::Int32; do_flips::Int32; umbrella::Int32; force_init::Int32; accum_mode::Int32
  # ABI v2 — mcmc_clustering_eap_chain.jl; all zero for this driver
  kappa::Cdouble; psi0::Cdouble; cutoff_radius::Cdouble; cluster_prob::Cdouble
  clustering::Int32; alpha_carry::Int32; cutoff_full::Int32; planar::Int32
end

pmc_error() = unsafe_string(ccall((:pmc_last_error, LIBPOLYMC), Cstring, ()))
check(rc) = rc == 0 || error("libpolymc_b200: $(pmc_error()) (status $rc)")

# (long, short, type, default) — one row per option of the reference's table
const OPTIONS = [
  ("--E0", "-e", Float64, 0.0), ("--chain-type", "-T", String, "dielectric"),
  ("--K1", "-J", Float64, 1.0), ("--K2", "-K", Float64, 0.0), ("--mu", "-m", Float64, 1e-2),
  ("--energy-type", "-u", String, "noninteracting"), ("--kT", "-k", Float64, 1.0),
  ("--ensemble-type", "-E", String, "force"), ("--Fz", "-F", Float64, 0.0), ("--Fx", "-G", Float64, 0.0),
  ("--rz", "-z", Float64, 0.0), ("--rx", "-x", Float64, 0.0), ("--mlen", "-b", Float64, 1.0),
  ("--num-monomers", "-n", Int, 100), ("--num-steps", "-N", Int, 100000), ("--num-inits", "-M", Int, 1),
  ("--phi-step", "-p", Float64, 3π/8), ("--theta-step", "-q", Float64, 3π/16),
  ("--chain-frac-step", "-f", Float64, 0.15), ("--step-adjust-lb", "-L", Float64, 0.15),
  ("--step-adjust-ub", "-U", Float64, 0.55), ("--step-adjust-scale", "-A", Float64, 1.1),
  ("--steps-per-adjust", "-S", Int, 2500), ("--acc", "-a", String, "metropolis"),
  ("--update-freq", nothing, Float64, 15.0), ("--verbose", "-v", Int, 3),
  ("--prefix", "-P", String, "eap-mcmc"), ("--postfix", "-Q", String, ""), ("--stepout", "-s", Int, 500),
  ("--numeric-type", nothing, String, "float64"),
  ("--replicas", nothing, Int, 1), ("--seed", nothing, Int, -1), ("--device", nothing, Int, 0),
  # [B200 path] number of GPUs of this box the replicas are spread over (pmc_multi_*, ABI v3); 0 = all of them
  ("--devices", nothing, Int, 1),
  # [B200 path] fp32: the rectangle of unchanged-dipole pairs of a trial in FP32 (pmc_set_pair_precision); else FP64
  ("--pair-precision", nothing, String, "fp64"),
]
const FLAGS = [("--force-init", "-I"), ("--do-flips", nothing), ("--umbrella-sampling", "-B"), ("--profile", "-Z")]

function cli()
  s = ArgParseSettings()
  for (long, short, T, dflt) in OPTIONS
    names = short === nothing ? long : [long, short]
    add_arg_table!(s, names, Dict(:arg_type => T, :default => dflt))
  end
  for (long, short) in FLAGS
    add_arg_table!(s, short === nothing ? long : [long, short], Dict(:action => :store_true))
  end
  return parse_args(s)
end

function case_of(p)
  ct = Dict("dielectric" => 0, "polar" => 1)
  et = Dict("noninteracting" => 0, "interacting" => 1, "Ising" => 2)
  haskey(ct, p["chain-type"]) || error("chain-type is not understood.")
  haskey(et, p["energy-type"]) || error("energy-type is not understood.")
  PmcCase(p["E0"], p["K1"], p["K2"], p["mu"], p["kT"], p["Fz"], p["Fx"], p["mlen"], p["phi-step"], p["theta-step"],
          p["step-adjust-lb"], p["step-adjust-ub"], p["step-adjust-scale"], p["num-monomers"], p["steps-per-adjust"],
          ct[p["chain-type"]], et[p["energy-type"]], p["do-flips"], p["umbrella-sampling"], p["force-init"],
          p["numeric-type"] == "float64" ? 0 : 1,
          0.0, 0.0, 7.5, 0.5, 0, 1, 0, 0)
end

# ConsoleLogger by --verbose, as mcmc_eap_chain.jl:157-165
function setup_logging(p)
  if p["verbose"] == 3
    global_logger(ConsoleLogger(stderr, Logging.Info))
  elseif p["verbose"] == 2
    global_logger(ConsoleLogger(stderr, Logging.Warn))
  elseif p["verbose"] == 1
    global_logger(ConsoleLogger(stderr, Logging.Error))
  else
    global_logger(Logging.NullLogger())
  end
end

# The hot loop in chunks, so that the progress log of mcmc_eap_chain.jl:294-299 stays alive on long runs: `run!` runs
# `todo` trials and returns the rows of chain 1 (8 × rows, 17 × rows) or nothing.
function run_init!(run!, p, nsteps, init, start, last_update, traj_io, rolling_io)
  stepout = p["stepout"]
  chunk = nsteps
  if nsteps > 200000
    chunk = stepout <= 0 ? 200000 : max(stepout, div(200000, stepout) * stepout)
  end
  done = 0
  while done < nsteps
    todo = min(chunk, nsteps - done)
    rows = run!(todo)
    if rows !== nothing
      writedlm(traj_io, permutedims(rows[1]), ',')
      writedlm(rolling_io, permutedims(rows[2]), ',')
    end
    done += todo
    if time() - last_update[] > p["update-freq"]
      @info "elapsed: $(time() - start)"
      @info "init:    $init / $(p["num-inits"])"
      @info "step:    $done / $nsteps"
      last_update[] = time()
    end
  end
end

function mcmc(nsteps::Int, p)
  p["acc"] == "metropolis" || error("'$(p["acc"])' acceptance criteria has not yet been implemented.")
  p["numeric-type"] in ("float64", "float128", "dec128", "big") || error("numeric-type '$(p["numeric-type"])' not understood")
  p["ensemble-type"] == "force" || error("ensemble-type '$(p["ensemble-type"])' is not supported by the B200 path (fixed-force only)")
  R = p["replicas"]; n = p["num-monomers"]; stepout = p["stepout"]
  seed = p["seed"] < 0 ? UInt64(time_ns()) & 0xffffffffffff : UInt64(p["seed"])
  cases = [case_of(p)]
  multi = p["devices"] != 1
  h = Ref{Ptr{Cvoid}}(C_NULL)
  if multi   # the R replicas over the GPUs of this box: one host thread per device inside the library (ABI v3)
    p["num-inits"] == 1 || error("--num-inits > 1 is per device handle: use --devices 1")
    check(ccall((:pmc_multi_create, LIBPOLYMC), Int32,
                (Ptr{PmcCase}, Int64, Int32, UInt64, Ptr{Int32}, Int32, Ptr{Ptr{Cvoid}}),
                cases, 1, R, seed, C_NULL, p["devices"], h))
  else
    check(ccall((:pmc_create, LIBPOLYMC), Int32, (Ptr{PmcCase}, Int64, Int32, UInt64, Int32, UInt32, Ptr{Ptr{Cvoid}}),
                cases, 1, R, seed, p["device"], 0, h))
  end
  prec = Dict("fp64" => 0, "fp32" => 1)
  haskey(prec, p["pair-precision"]) || error("pair-precision is not understood.")
  if multi
    check(ccall((:pmc_multi_set_pair_precision, LIBPOLYMC), Int32, (Ptr{Cvoid}, Int32), h[], prec[p["pair-precision"]]))
  else
    check(ccall((:pmc_set_pair_precision, LIBPOLYMC), Int32, (Ptr{Cvoid}, Int32), h[], prec[p["pair-precision"]]))
  end
  traj_io = open("$(p["prefix"])_trajectory.csv", "w"); rolling_io = open("$(p["prefix"])_rolling.csv", "w")
  writedlm(traj_io, ["step" "r1" "r2" "r3" "p1" "p2" "p3" "U"], ',')
  writedlm(rolling_io, permutedims(vcat(["step"], ["r1","r2","r3","r1sq","r2sq","r3sq","rsq","p1","p2","p3","p1sq","p2sq","p3sq","psq","U","Usq"])), ',')
  start = time(); last_update = Ref(start)
  try
    function run!(todo)
      rows = multi ? ccall((:pmc_multi_rows_for, LIBPOLYMC), Int64, (Ptr{Cvoid}, Int64, Int64), h[], todo, stepout) :
                     ccall((:pmc_rows_for, LIBPOLYMC), Int64, (Ptr{Cvoid}, Int64, Int64), h[], todo, stepout)
      traj = Array{Float64}(undef, 8, rows, R); roll = Array{Float64}(undef, 17, rows, R)   # C order [chain][row][k]
      if multi
        check(ccall((:pmc_multi_run, LIBPOLYMC), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                    h[], todo, stepout, traj, roll))
      else
        check(ccall((:pmc_run, LIBPOLYMC), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                    h[], todo, stepout, traj, roll))
      end
      return rows > 0 ? (traj[:, :, 1], roll[:, :, 1]) : nothing
    end
    for init in 1:p["num-inits"]
      run_init!(run!, p, nsteps, init, start, last_update, traj_io, rolling_io)
      init < p["num-inits"] && check(ccall((:pmc_reinit, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Int32}), h[], C_NULL))
    end
    if multi    # one ncclAllGather of the [R][24] result rows inside the library
      table = Array{Float64}(undef, 24, R)
      check(ccall((:pmc_multi_gather, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], table))
      nrm = table[18, :]
      avg = [sum(table[k, :] .* nrm) for k in 1:16] ./ sum(nrm)      # pooled Σ values / Σ normalisers
      ar = sum(table[17, :] .* table[21, :]) / (R * p["num-steps"])
    else
      sums = Array{Float64}(undef, 17, R); diag = Array{Float64}(undef, 8, R)
      check(ccall((:pmc_accumulators, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], sums))
      check(ccall((:pmc_diagnostics, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], diag))
      if p["umbrella-sampling"]   # replicas carry different gauges exp(Ω0): pool their ratios (see polymc/mcmc.py)
        avg = vec(sum(sums[1:16, :] ./ sums[17:17, :], dims=2)) ./ R
      else
        pooled = vec(sum(sums, dims=2)); avg = pooled[1:16] ./ pooled[17]
      end
      ar = sum(diag[5, :]) / (R * p["num-inits"] * p["num-steps"])
    end
    @info "total time elapsed: $(time() - start)"
    @info "acceptance rate: $ar"
    return avg, ar
  finally
    close(traj_io); close(rolling_io)
    if multi
      ccall((:pmc_multi_destroy, LIBPOLYMC), Cvoid, (Ptr{Cvoid},), h[])
    else
      ccall((:pmc_destroy, LIBPOLYMC), Cvoid, (Ptr{Cvoid},), h[])
    end
  end
end

function main()
  p = cli()
  setup_logging(p)
  p["profile"] && error("not implemented for the HPC env")
  avg, ar = mcmc(p["num-steps"], p)
  nb = p["mlen"] * p["num-monomers"]
  println("<r>    =   $(avg[1:3])");   println("<r/nb> =   $(avg[1:3] / nb)")
  println("<rj2>  =   $(avg[4:6])");   println("<r2>   =   $(avg[7])")
  println("<p>    =   $(avg[8:10])");  println("<pj2>  =   $(avg[11:13])")
  println("<p2>   =   $(avg[14])");    println("<U>    =   $(avg[15])")
  println("<U2>   =   $(avg[16])");    println("AR     =   $ar")
end

abspath(PROGRAM_FILE) == @__FILE__ && main()
