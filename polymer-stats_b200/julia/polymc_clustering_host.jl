# polymc_clustering_host.jl — Julia host for the clustering driver on libpolymc_b200.so (ABI v2).
#
# Drop-in for `julia mcmc_clustering_eap_chain.jl ...`: same command line (ArgParse names/defaults of
# mcmc_clustering_eap_chain.jl:19-153), same 12 stdout lines (:389-400), same two CSV files (:253-259);
# the two `mcmc(...)` methods (:166-352) and the burn-in ladder (:365-386) become `ccall`s:
#   pmc_create [+ pmc_init_x0]  →  for each kT multiplier: pmc_begin_stage, pmc_run_ex  →  pmc_begin_stage(1),
#   pmc_run_ex with rows  →  pmc_accumulators, pmc_extra_accumulators, pmc_diagnostics.
#
# NOTE: no Julia runtime in the build image; tools/minijl executes this file against the real library (ccall through
# ctypes) and tests/test_gpu_julia_hosts.py compares stdout and both CSV files with the Python twin byte for byte.
using ArgParse, Printf, DelimitedFiles
include(joinpath(@__DIR__, "polymc_host.jl"))   # PmcCase, check, LIBPOLYMC

const CL_OPTIONS = [
  ("--E0", "-e", Float64, 0.0), ("--chain-type", "-T", String, "dielectric"),
  ("--K1", "-J", Float64, 1.0), ("--K2", "-K", Float64, 0.0), ("--mu", "-m", Float64, 1e-2),
  ("--bend-mod", "-a", Float64, 0.0), ("--bend-angle", "-g", Float64, 0.0),
  ("--energy-type", "-u", String, "Ising"), ("--cutoff-radius", nothing, Float64, 7.5),
  ("--kT", "-k", Float64, 1.0), ("--Fz", "-F", Float64, 0.0), ("--Fx", "-G", Float64, 0.0),
  ("--mlen", "-b", Float64, 1.0), ("--num-monomers", "-n", Int, 100), ("--num-steps", "-N", Int, 1000000),
  ("--phi-step", "-p", Float64, 3π/8), ("--theta-step", "-q", Float64, 3π/16), ("--cluster-prob", nothing, Float64, 0.5),
  ("--step-adjust-lb", "-L", Float64, 0.15), ("--step-adjust-ub", "-U", Float64, 0.40),
  ("--step-adjust-scale", "-A", Float64, 1.1), ("--steps-per-adjust", "-S", Int, 2500),
  ("--update-freq", nothing, Float64, 15.0), ("--verbose", "-v", Int, 3),
  ("--prefix", "-P", String, "eap-mcmc"), ("--postfix", "-Q", String, ""), ("--stepout", "-s", Int, 500),
  ("--numeric-type", nothing, String, "float64"), ("--burn-in", nothing, Int, 50000),
  ("--burn-schedule", nothing, String, "[1000; 100; 10; 2; 1]"), ("--x0", nothing, String, nothing),
  ("--dx0", nothing, String, "[2*pi, 1e-1]"),
  ("--replicas", nothing, Int, 1), ("--seed", nothing, Int, -1), ("--device", nothing, Int, 0),
]
const CL_FLAGS = [("--umbrella-sampling", "-B"), ("--profile", "-Z"), ("--no-alpha-carry", nothing),
                  ("--cutoff-full-energy", nothing)]

function cl_cli()
  s = ArgParseSettings()
  for (long, short, T, dflt) in CL_OPTIONS
    names = short === nothing ? long : [long, short]
    add_arg_table!(s, names, dflt === nothing ? Dict(:arg_type => T) : Dict(:arg_type => T, :default => dflt))
  end
  for (long, short) in CL_FLAGS
    add_arg_table!(s, short === nothing ? long : [long, short], Dict(:action => :store_true))
  end
  return parse_args(s)
end

function cl_case_of(p)
  ct = Dict("dielectric" => 0, "polar" => 1)
  et = Dict("noninteracting" => 0, "interacting" => 1, "Ising" => 2, "cutoff" => 3)
  haskey(ct, p["chain-type"]) || error("chain-type is not understood.")
  haskey(et, p["energy-type"]) || error("energy-type is not understood.")
  PmcCase(p["E0"], p["K1"], p["K2"], p["mu"], p["kT"], p["Fz"], p["Fx"], p["mlen"], p["phi-step"], p["theta-step"],
          p["step-adjust-lb"], p["step-adjust-ub"], p["step-adjust-scale"], p["num-monomers"], p["steps-per-adjust"],
          ct[p["chain-type"]], et[p["energy-type"]], 0, p["umbrella-sampling"], 0, p["numeric-type"] == "float64" ? 0 : 1,
          p["bend-mod"], p["bend-angle"], p["cutoff-radius"], p["cluster-prob"], 1, p["no-alpha-carry"] ? 0 : 1,
          p["cutoff-full-energy"] ? 1 : 0, 0)
end

# μ columns of the trajectory file (:318) from the dumped angles (inc/dipole_response.jl:7-29)
function dipoles(p, ϕ, θ)
  n̂ = [cos(ϕ)*sin(θ), sin(ϕ)*sin(θ), cos(θ)]
  p["chain-type"] == "dielectric" ? (p["K1"]-p["K2"])*p["E0"]*cos(θ)*n̂ + [0.0, 0.0, p["K2"]*p["E0"]] : p["mu"]*n̂
end

const T_START = Ref(time())
const T_LAST = Ref(time())

# One `mcmc(nsteps, pargs, chain)` call of the reference (:171-352) = pmc_begin_stage + the loop.  The loop runs in
# chunks of whole --stepout intervals so that the progress log of :289-293 stays alive on long stages.
function stage!(h, nsteps, p, kT_scale, write_files)
  R = p["replicas"]; n = p["num-monomers"]; stepout = write_files ? p["stepout"] : 0
  check(ccall((:pmc_begin_stage, LIBPOLYMC), Int32, (Ptr{Cvoid}, Cdouble), h, kT_scale))   # mcmc(nsteps, pargs, chain) :171-265
  traj_io = write_files ? open("$(p["prefix"])_trajectory.csv", "w") : nothing
  roll_io = write_files ? open("$(p["prefix"])_rolling.csv", "w") : nothing
  if write_files
    writedlm(traj_io, hcat(["step" "r1" "r2" "r3" "p1" "p2" "p3" "U"],
                           reshape(vcat(reshape(["phi$i" for i=1:n], 1, :), reshape(["theta$i" for i=1:n], 1, :)), 1, :),
                           reshape(vcat(reshape(["mux$i" for i=1:n], 1, :), reshape(["muy$i" for i=1:n], 1, :), reshape(["muz$i" for i=1:n], 1, :)), 1, :)), ',')
    writedlm(roll_io, ["step" "r1" "r2" "r3" "r1sq" "r2sq" "r3sq" "rsq" "p1" "p2" "p3" "p1sq" "p2sq" "p3sq" "psq" "U" "Usq" "Ealign" "psi"], ',')
  end
  chunk = nsteps <= 200000 ? nsteps : (stepout <= 0 ? 200000 : max(stepout, div(200000, stepout) * stepout))
  done = 0
  try
    while done < nsteps
      todo = min(chunk, nsteps - done)
      rows = ccall((:pmc_rows_for, LIBPOLYMC), Int64, (Ptr{Cvoid}, Int64, Int64), h, todo, stepout)
      traj = Array{Float64}(undef, 8, rows, R); roll = Array{Float64}(undef, 19, rows, R); state = Array{Float64}(undef, 2n, rows, R)
      check(ccall((:pmc_run_ex, LIBPOLYMC), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                  h, todo, stepout, rows > 0 ? traj : C_NULL, rows > 0 ? roll : C_NULL, rows > 0 ? state : C_NULL))  # loop :267-336
      if write_files
        for r in 1:rows
          μs = hcat([dipoles(p, state[2i-1, r, 1], state[2i, r, 1]) for i in 1:n]...)
          writedlm(traj_io, hcat(transpose(traj[:, r, 1]), transpose(state[:, r, 1]), reshape(μs, 1, :)), ',')
        end
        rows > 0 && writedlm(roll_io, permutedims(roll[:, :, 1]), ',')
      end
      done += todo
      if time() - T_LAST[] > p["update-freq"]                                                # :289-293
        @info "elapsed: $(time() - T_START[])"
        @info "step:    $done / $nsteps"
        T_LAST[] = time()
      end
    end
  finally
    write_files && (close(traj_io); close(roll_io))
  end
  @info "total time elapsed: $(time() - T_START[])"                                          # :334-335
end

function cl_main()
  p = cl_cli()
  setup_logging(p)                                                       # ConsoleLogger by --verbose, :157-165
  p["profile"] && error("Not currently implemented...")
  p["numeric-type"] in ("float64", "float128", "dec128", "big") || error("numeric-type '$(p["numeric-type"])' not understood")
  R = p["replicas"]
  seed = p["seed"] < 0 ? UInt64(time_ns()) & 0xffffffffffff : UInt64(p["seed"])
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:pmc_create, LIBPOLYMC), Int32, (Ptr{PmcCase}, Int64, Int32, UInt64, Int32, UInt32, Ptr{Ptr{Cvoid}}),
              [cl_case_of(p)], 1, R, seed, p["device"], 0, h))
  try
    if p["x0"] !== nothing                                               # inc/eap_chain.jl:63-78
      x0 = Float64.(eval(Meta.parse(p["x0"]))); dx0 = Float64.(eval(Meta.parse(p["dx0"])))
      check(ccall((:pmc_init_x0, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), h[], x0, length(x0), dx0))
    end
    for mult in eval(Meta.parse(p["burn-schedule"]))                     # :365-381
      stage!(h[], p["burn-in"], p, Float64(mult), false)
    end
    stage!(h[], p["num-steps"], p, 1.0, true)                            # :383-384
    sums = Array{Float64}(undef, 17, R); xs = Array{Float64}(undef, 2, R); diag = Array{Float64}(undef, 8, R)
    check(ccall((:pmc_accumulators, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], sums))
    check(ccall((:pmc_extra_accumulators, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], xs))
    check(ccall((:pmc_diagnostics, LIBPOLYMC), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], diag))
    if p["umbrella-sampling"]     # replicas carry different gauges exp(Ω0): pool their ratios (polymc/mcmc.py pool_replicas)
      avg = vec(sum(sums[1:16, :] ./ sums[17:17, :], dims=2)) ./ R; ex = vec(sum(xs ./ sums[17:17, :], dims=2)) ./ R
    else
      pooled = vec(sum(sums, dims=2)); avg = pooled[1:16] ./ pooled[17]; ex = vec(sum(xs, dims=2)) ./ pooled[17]
    end
    ar = sum(diag[5, :]) / (R * p["num-steps"])
    @info "acceptance rate: $ar"
    nb = p["mlen"] * p["num-monomers"]
    println("<r>    =   $(avg[1:3])");   println("<r/nb> =   $(avg[1:3] / nb)")
    println("<rj2>  =   $(avg[4:6])");   println("<r2>   =   $(avg[7])")
    println("<p>    =   $(avg[8:10])");  println("<pj2>  =   $(avg[11:13])")
    println("<p2>   =   $(avg[14])");    println("<U>    =   $(avg[15])")
    println("<U2>   =   $(avg[16])");    println("<cos2(θ)>   =   $(ex[1])")
    println("<ψ>    =   $(ex[2])");      println("AR     =   $ar")
  finally
    ccall((:pmc_destroy, LIBPOLYMC), Cvoid, (Ptr{Cvoid},), h[])
  end
end

abspath(PROGRAM_FILE) == @__FILE__ && cl_main()
