# polymc_clustering_2d_host.jl — Julia host for the 2-D tree's driver on libpolymc_b200.so (ABI v2, planar = 1).
#
# Drop-in for `julia 2D/mcmc_clustering_eap_chain.jl ...`: same command line (ArgParse names/defaults of
# 2D/mcmc_clustering_eap_chain.jl:19-133), same 10 stdout lines with 2-vectors (:338-347), same two CSV files
# (6-column trajectory :228, 13-column rolling :230); `mcmc(...)` (:143-311) and the burn-in ladder (:323-336)
# become `ccall`s:
#   pmc_create(planar = 1)  →  for each kT multiplier: pmc_begin_stage, pmc_run  →  pmc_begin_stage(1),
#   pmc_run with rows  →  pmc_accumulators, pmc_diagnostics.
# The planar chain is the 3-D chain restricted to the x–z plane (state ϕ only, n̂ = (cosϕ, sinϕ), field along the
# second axis: 2D/inc/eap_chain.jl:33, 2D/inc/dipole_response.jl:7-10), so the library returns 3-D rows and this
# host keeps the x ("1") and z ("3") columns.  As upstream, every `mcmc()` call starts from a NEW random chain
# (`chain = EAPChain(pargs)`, :151): `pmc_begin_stage` on a planar handle redraws the chain.
#
# NOTE: no Julia runtime in the build image; tools/minijl executes this file against the real library (ccall through
# ctypes) and tests/test_gpu_julia_hosts.py compares stdout and both CSV files with the Python twin byte for byte.
using ArgParse, Printf, DelimitedFiles
include(joinpath(@__DIR__, "polymc_host.jl"))   # PmcCase, check, LIBPOLYMC

const CL2D_OPTIONS = [
  ("--E0", "-e", Float64, 0.0), ("--chain-type", "-T", String, "dielectric"),
  ("--K1", "-J", Float64, 1.0), ("--K2", "-K", Float64, 0.0), ("--mu", "-m", Float64, 1e-2),
  ("--energy-type", "-u", String, "noninteracting"),
  ("--kT", "-k", Float64, 1.0), ("--Fz", "-F", Float64, 0.0), ("--Fx", "-G", Float64, 0.0),
  ("--mlen", "-b", Float64, 1.0), ("--num-monomers", "-n", Int, 100), ("--num-steps", "-N", Int, 1000000),
  ("--phi-step", "-p", Float64, 3π/8), ("--cluster-prob", nothing, Float64, 0.5),
  ("--step-adjust-lb", "-L", Float64, 0.15), ("--step-adjust-ub", "-U", Float64, 0.40),
  ("--step-adjust-scale", "-A", Float64, 1.1), ("--steps-per-adjust", "-S", Int, 2500),
  ("--update-freq", nothing, Float64, 15.0), ("--verbose", "-v", Int, 3),
  ("--prefix", "-P", String, "eap-mcmc"), ("--postfix", "-Q", String, ""), ("--stepout", "-s", Int, 500),
  ("--numeric-type", nothing, String, "float64"), ("--burn-in", nothing, Int, 50000),
  ("--burn-schedule", nothing, String, "[1000; 100; 10; 2; 1]"),
  ("--replicas", nothing, Int, 1), ("--seed", nothing, Int, -1), ("--device", nothing, Int, 0),
]
const CL2D_FLAGS = [("--umbrella-sampling", "-B"), ("--profile", "-Z"), ("--no-alpha-carry", nothing)]

function cl2d_cli()
  s = ArgParseSettings()
  for (long, short, T, dflt) in CL2D_OPTIONS
    add_arg_table!(s, short === nothing ? long : [long, short], Dict(:arg_type => T, :default => dflt))
  end
  for (long, short) in CL2D_FLAGS
    add_arg_table!(s, short === nothing ? long : [long, short], Dict(:action => :store_true))
  end
  return parse_args(s)
end

function cl2d_case_of(p)
  ct = Dict("dielectric" => 0, "polar" => 1)
  et = Dict("noninteracting" => 0, "interacting" => 1, "Ising" => 2)          # no cut-off energy in the 2-D tree
  haskey(ct, p["chain-type"]) || error("chain-type is not understood.")      # 2D/inc/eap_chain.jl:76-78
  haskey(et, p["energy-type"]) || error("energy-type is not understood.")    # 2D/inc/eap_chain.jl:88-90
  # There is no θ: θstep is tied to ϕstep/2 so that the shared adaptation rule caps at ϕstep = π exactly as
  # `ϕstep != π` does (2D/mcmc_clustering_eap_chain.jl:262-273).
  PmcCase(p["E0"], p["K1"], p["K2"], p["mu"], p["kT"], p["Fz"], p["Fx"], p["mlen"], p["phi-step"], p["phi-step"] / 2,
          p["step-adjust-lb"], p["step-adjust-ub"], p["step-adjust-scale"], p["num-monomers"], p["steps-per-adjust"],
          ct[p["chain-type"]], et[p["energy-type"]], 0, p["umbrella-sampling"], 0, p["numeric-type"] == "float64" ? 0 : 1,
          0.0, 0.0, 7.5, p["cluster-prob"], 1, p["no-alpha-carry"] ? 0 : 1, 0, 1)
end

const TRAJ_COLS_2D = [1, 2, 4, 5, 7, 8]                                    # step, r1, r3, p1, p3, U of the 8 columns
const ROLL_COLS_2D = [1, 2, 4, 5, 7, 8, 9, 11, 12, 14, 15, 16, 17]         # of the 17 rolling columns

function stage2d!(h, nsteps, p, kT_scale, write_files)
  R = p["replicas"]; stepout = write_files ? p["stepout"] : 0
  check(ccall((:pmc_begin_stage, LIBPOLYMC), Int32, (Ptr{Cvoid}, Cdouble), h, kT_scale))   # a fresh mcmc(...) call, :143-151
  rows = ccall((:pmc_rows_for, LIBPOLYMC), Int64, (Ptr{Cvoid}, Int64, Int64), h, nsteps, stepout)
  traj = Array{Float64}(undef, 8, rows, R); roll = Array{Float64}(undef, 17, rows, R)
  t0 = time()
  check(ccall((:pmc_run, LIBPOLYMC), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
              h, nsteps, stepout, rows > 0 ? traj : C_NULL, rows > 0 ? roll : C_NULL))     # loop :236-300
  @info "step:    $nsteps / $nsteps"                                                        # :254-258
  @info "total time elapsed: $(time() - t0)"                                               # :302-303
  write_files || return
  open("$(p["prefix"])_trajectory.csv", "w") do io
    writedlm(io, ["step" "r1" "r3" "p1" "p3" "U"], ',')                                    # :228
    rows > 0 && writedlm(io, permutedims(traj[TRAJ_COLS_2D, :, 1]), ',')
  end
  open("$(p["prefix"])_rolling.csv", "w") do io
    writedlm(io, ["step" "r1" "r3" "r1sq" "r3sq" "rsq" "p1" "p3" "p1sq" "p3sq" "psq" "U" "Usq"], ',')   # :230
    rows > 0 && writedlm(io, permutedims(roll[ROLL_COLS_2D, :, 1]), ',')
  end
end

function cl2d_main()
  p = cl2d_cli()
  setup_logging(p)                                                        # ConsoleLogger by --verbose
  p["profile"] && error("Not currently implemented...")
  p["numeric-type"] in ("float64", "float128", "dec128", "big") || error("numeric-type '$(p["numeric-type"])' not understood")
  R = p["replicas"]
  seed = p["seed"] < 0 ? UInt64(time_ns()) & 0xffffffffffff : UInt64(p["seed"])
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:pmc_create, LIBPOLYMC), Int32, (Ptr{PmcCase}, Int64, Int32, UInt64, Int32, UInt32, Ptr{Ptr{Cvoid}}),
              [cl2d_case_of(p)], 1, R, seed, p["device"], 0, h))
  try
    for mult in eval(Meta.parse(p["burn-schedule"]))                      # :323-333
      stage2d!(h[], p["burn-in"], p, Float64(mult), false)
    end
    stage2d!(h[], p["num-steps"], p, 1.0, true)                           # :335-336
    sums = Array{Float64}(undef, 17, R); diag = Array{Float64}(undef, 8, R)
    check(ccall((:pmc_accumulators, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], sums))
    check(ccall((:pmc_diagnostics, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], diag))
    if p["umbrella-sampling"]     # replicas carry different gauges: pool their ratios (polymc/mcmc.py pool_replicas)
      avg = vec(sum(sums[1:16, :] ./ sums[17:17, :], dims=2)) ./ R
    else
      pooled = vec(sum(sums, dims=2)); avg = pooled[1:16] ./ pooled[17]
    end
    ar = sum(diag[5, :]) / (R * p["num-steps"])
    @info "acceptance rate: $ar"
    nb = p["mlen"] * p["num-monomers"]
    xz(v) = [v[1], v[3]]                                                   # the plane of the chain
    println("<r>    =   $(xz(avg[1:3]))");   println("<r/nb> =   $(xz(avg[1:3]) / nb)")
    println("<rj2>  =   $(xz(avg[4:6]))");   println("<r2>   =   $(avg[7])")
    println("<p>    =   $(xz(avg[8:10]))");  println("<pj2>  =   $(xz(avg[11:13]))")
    println("<p2>   =   $(avg[14])");        println("<U>    =   $(avg[15])")
    println("<U2>   =   $(avg[16])");        println("AR     =   $ar")
  finally
    ccall((:pmc_destroy, LIBPOLYMC), Cvoid, (Ptr{Cvoid},), h[])
  end
end

abspath(PROGRAM_FILE) == @__FILE__ && cl2d_main()
