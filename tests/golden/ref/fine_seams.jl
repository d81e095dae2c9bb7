# fine_seams.jl — the reference's own energy / move / cluster / acceptance functions on fixed configurations.
#
#   julia tests/golden/ref/fine_seams.jl <reference root> <tape.txt> <out.json>
#
# Includes the UNMODIFIED inc/eap_chain.jl (+ dipole_response.jl, energy.jl), inc/average.jl and inc/acceptance.jl of
# the reference and writes, for a set of chains: U(chain) (inc/eap_chain.jl:411 → inc/energy.jl:7-23), U_interaction
# (:196-211), U_Ising (:215-228), UCutoff (:171-192), Ω, r, p, ψ; `move!` (:230-257) on copies; `cluster_flip!`
# (:269-333) with its α; the Metropolis functor (inc/acceptance.jl:29-37) incl. the α carry; metropolis_acc (:1-3);
# AntiDipoleWeightFunction (inc/average.jl:104-124).  tests/test_reference_pin.py checks the CPU oracle against this
# file and tests/test_gpu_reference_pin.py the CUDA library.
#
# Valid Julia; executed here by tools/minijl (no Julia in the image) — see prelude.jl.
using Distributions
using LinearAlgebra
include(joinpath(@__DIR__, "prelude.jl"))

const REF_ROOT = ARGS[1]
load_tape!(ARGS[2])
const OUT_PATH = ARGS[3]

include(joinpath(REF_ROOT, "inc", "eap_chain.jl"))
include(joinpath(REF_ROOT, "inc", "average.jl"))
include(joinpath(REF_ROOT, "inc", "acceptance.jl"))

# ---- a minimal JSON writer (no packages) ----
jnum(x::Bool) = x ? "true" : "false"
jnum(x::Integer) = string(x)
function jnum(x::AbstractFloat)
  if isnan(x)
    return "\"NaN\""
  elseif isinf(x)
    return x > 0 ? "\"Inf\"" : "\"-Inf\""
  end
  return repr(x)
end
jvec(v) = "[" * join(map(jnum, v), ",") * "]"
jstr(s) = "\"" * s * "\""
jobj(pairs) = "{" * join(map(p -> jstr(p[1]) * ":" * p[2], pairs), ",") * "}"
jlist(items) = "[" * join(items, ",") * "]"

function make_pargs(ct, et, nmono, E0, K1, K2, mu, kT, Fz, Fx, b, kappa, psi0, crad)
  pargs = Dict{String,Any}()
  pargs["num-monomers"] = nmono
  pargs["mlen"] = b
  pargs["E0"] = E0
  pargs["K1"] = K1
  pargs["K2"] = K2
  pargs["mu"] = mu
  pargs["kT"] = kT
  pargs["Fz"] = Fz
  pargs["Fx"] = Fx
  pargs["chain-type"] = ct
  pargs["energy-type"] = et
  pargs["bend-mod"] = kappa
  pargs["bend-angle"] = psi0
  pargs["cutoff-radius"] = crad
  return pargs
end

function chain_record(chain)
  return [("U", jnum(chain.U)), ("Omega", jnum(chain.Ω)), ("r", jvec(chain.r)), ("p", jvec(chain_μ(chain))),
          ("sum_us", jnum(sum(chain.us))), ("U_interaction", jnum(U_interaction(chain))),
          ("U_Ising", jnum(U_Ising(chain))), ("U_cutoff", jnum(UCutoff(chain.b * 2.0)(chain))),
          ("sum_psi", jnum(sum(chain.ψs))), ("sum_cos2", jnum(sum(map(x -> x*x, chain.cθs)))),
          ("xs_last", jvec(chain.xs[:, end]))]
end

function one_case(name, pargs, nmoves, nflips)
  nmono = pargs["num-monomers"]
  chain = EAPChain(pargs)          # random init: 2n uniforms from the tape (eap_chain.jl:62)
  chain.U = U(chain)
  fields = [("name", jstr(name)), ("chain_type", jstr(pargs["chain-type"])), ("energy_type", jstr(pargs["energy-type"])),
            ("n", jnum(nmono)), ("b", jnum(pargs["mlen"])), ("E0", jnum(pargs["E0"])), ("K1", jnum(pargs["K1"])),
            ("K2", jnum(pargs["K2"])), ("mu", jnum(pargs["mu"])), ("kT", jnum(pargs["kT"])), ("Fz", jnum(pargs["Fz"])),
            ("Fx", jnum(pargs["Fx"])), ("kappa", jnum(pargs["bend-mod"])), ("psi0", jnum(pargs["bend-angle"])),
            ("cutoff_radius", jnum(pargs["cutoff-radius"])),
            ("phi", jvec(chain.ϕs)), ("theta", jvec(chain.θs))]
  append!(fields, chain_record(chain))

  # AntiDipoleWeightFunction (average.jl:104-124) and logπ (acceptance.jl:19-23)
  wf = AntiDipoleWeightFunction(chain)
  push!(fields, ("log_gauge", jnum(wf.log_gauge)))
  push!(fields, ("weight", jnum(wf(chain))))
  push!(fields, ("logpi_weightless", jnum(logπ_chain(chain, WeightlessFunction()))))
  push!(fields, ("logpi_umbrella", jnum(logπ_chain(chain, wf))))

  # move! on copies (eap_chain.jl:230-257), incl. θ clamped at both ends and the flip of --do-flips
  moves = Any[]
  for k in 1:nmoves
    idx = rand(1:nmono)
    dϕ = rand(Uniform(-1.2, 1.2))
    dθ = rand(Uniform(-0.6, 0.6))
    if k == nmoves - 2
      dθ = 5.0                       # clamps to π
    elseif k == nmoves - 1
      dθ = -5.0                      # clamps to 0: sinθ = 0, Ω = -Inf
    elseif k == nmoves
      dθ = π - 2*chain.θs[idx]       # the flip of mcmc_eap_chain.jl:279
      dϕ = π
    end
    t = EAPChain(chain)
    move!(t, idx, dϕ, dθ)
    push!(moves, jobj([("idx", jnum(idx)), ("dphi", jnum(dϕ)), ("dtheta", jnum(dθ)), ("U", jnum(t.U)),
                       ("Omega", jnum(t.Ω)), ("r", jvec(t.r)), ("p", jvec(chain_μ(t))), ("sum_us", jnum(sum(t.us))),
                       ("sum_psi", jnum(sum(t.ψs))), ("theta_new", jnum(t.θs[idx]))]))
  end
  push!(fields, ("moves", jlist(moves)))

  # move! + cluster_flip! (eap_chain.jl:269-333) on copies; the tape decides gate and growth
  flips = Any[]
  for k in 1:nflips
    idx = rand(1:nmono)
    dϕ = rand(Uniform(-1.0, 1.0))
    dθ = rand(Uniform(-0.5, 0.5))
    t = EAPChain(chain)
    move!(t, idx, dϕ, dθ)
    θ_moved = copy(t.θs)
    U_moved = t.U
    used0 = RAND_USED[1]
    α = cluster_flip!(t, idx; ϵflip = 0.25)
    changed = [i for i in 1:nmono if t.θs[i] != θ_moved[i]]
    lo = length(changed) > 0 ? minimum(changed) : 0
    hi = length(changed) > 0 ? maximum(changed) : 0
    push!(flips, jobj([("idx", jnum(idx)), ("dphi", jnum(dϕ)), ("dtheta", jnum(dθ)), ("alpha", jnum(α)),
                       ("lo", jnum(lo)), ("hi", jnum(hi)), ("draws", jnum(RAND_USED[1] - used0)),
                       ("U_moved", jnum(U_moved)), ("U", jnum(t.U)), ("Omega", jnum(t.Ω)), ("r", jvec(t.r)),
                       ("p", jvec(chain_μ(t))), ("sum_us", jnum(sum(t.us))), ("sum_psi", jnum(sum(t.ψs))),
                       ("sum_cos2", jnum(sum(map(x -> x*x, t.cθs)))), ("theta", jvec(t.θs)), ("phi", jvec(t.ϕs))]))
  end
  push!(fields, ("cluster_flips", jlist(flips)))

  # the Metropolis functor with the α carry (acceptance.jl:24-37): a scripted sequence of trials
  acc = Metropolis(chain, WeightlessFunction())
  cur = chain
  seq = Any[]
  for k in 1:12
    idx = rand(1:nmono)
    t = EAPChain(cur)
    move!(t, idx, rand(Uniform(-0.8, 0.8)), rand(Uniform(-0.4, 0.4)))
    α = (k % 3 == 0) ? 0.5 + rand() : 1.0
    ϵ = rand()
    prev = acc.logπ_prev
    ok = acc(t, ϵ; α = α)
    if ok
      cur = t
    end
    push!(seq, jobj([("idx", jnum(idx)), ("alpha", jnum(α)), ("eps", jnum(ϵ)), ("logpi_prev_before", jnum(prev)),
                     ("logpi_trial", jnum(logπ_chain(t, WeightlessFunction()))), ("accepted", jnum(ok)),
                     ("logpi_prev_after", jnum(acc.logπ_prev)), ("U_current", jnum(cur.U))]))
  end
  push!(fields, ("metropolis", jlist(seq)))
  return jobj(fields)
end

function main()
  cases = Any[]
  k = 0
  for ct in ["dielectric", "polar"]
    for et in ["noninteracting", "interacting", "Ising", "cutoff"]
      for nmono in [6, 13]
        k += 1
        kappa = (k % 2 == 0) ? 0.5 : 0.0
        pargs = make_pargs(ct, et, nmono, 1.0 + 0.25*k, 1.0, 0.25, 0.2 + 0.05*k, 0.8 + 0.1*k, 0.7 - 0.1*k, 0.3, 1.0 + 0.05*k,
                           kappa, 0.3, 2.0)
        push!(cases, one_case("case$k", pargs, 8, 6))
      end
    end
  end
  # metropolis_acc (acceptance.jl:1-3), the re-initialisation rule of mcmc_eap_chain.jl:352-357
  macc = Any[]
  for k in 1:8
    kT = 0.5 + rand()
    dU = 4.0 * (rand() - 0.5)
    sa = 0.1 + rand()
    sb = 0.1 + rand()
    ϵ = rand()
    push!(macc, jobj([("kT", jnum(kT)), ("dU", jnum(dU)), ("s_a", jnum(sa)), ("s_b", jnum(sb)), ("eps", jnum(ϵ)),
                      ("accept", jnum(metropolis_acc(kT, dU, sa, sb, ϵ)))]))
  end
  io = open(OUT_PATH, "w")
  write(io, jobj([("generator", jstr("tests/golden/ref/fine_seams.jl on the unmodified reference sources")),
                  ("cases", jlist(cases)), ("metropolis_acc", jlist(macc)), ("tape_used", jnum(RAND_USED[1]))]))
  write(io, "\n")
  close(io)
end
main()
