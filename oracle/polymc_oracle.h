/*
 * polymc_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the fixed-force MCMC hot path of grasingerm/polymer-stats
 * (mcmc_eap_chain.jl + inc/eap_chain.jl, energy.jl, dipole_response.jl, acceptance.jl,
 * average.jl) and of its clustering twin mcmc_clustering_eap_chain.jl (cluster_flip!, bending
 * energy, UCutoff, burn-in stages).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load this library; the shipped CUDA library never does.
 *
 * PARITY UNPINNED by the reference: the reference ships no tests, no golden vectors and no
 * result data, and Julia is not installed in this image, so the reference cannot be run here.
 * The oracle is pinned instead by (1) the known-answer values of an independent numpy
 * restatement recorded in SURVEY.md §8c (tests/golden/kat_n5.json), (2) closed-form
 * single-monomer integrals for non-interacting chains (oracle/closed_form.py) and (3) the
 * agreement of its two internal formulations (full recompute, as the reference does, versus
 * the changed-pair ΔU the CUDA path uses).
 *
 * Every function cites the reference file:line it restates.
 */
#ifndef POLYMC_ORACLE_H
#define POLYMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_CHAIN_DIELECTRIC = 0, ORC_CHAIN_POLAR = 1 };
enum { ORC_ENERGY_NONINTERACTING = 0, ORC_ENERGY_INTERACTING = 1, ORC_ENERGY_ISING = 2, ORC_ENERGY_CUTOFF = 3 };

/* One case = one command line of mcmc_eap_chain.jl (mcmc_eap_chain.jl:19-153). */
typedef struct orc_case {
  double E0, K1, K2, mu, kT, Fz, Fx, b;
  double phi_step, theta_step;
  double adj_lb, adj_ub, adj_scale;
  int64_t n;
  int64_t steps_per_adjust;
  int32_t chain_type;   /* ORC_CHAIN_*  */
  int32_t energy_type;  /* ORC_ENERGY_* */
  int32_t do_flips;
  int32_t umbrella;
  int32_t omega_compat; /* 1: Omega0 = log(prod(sin theta)) exactly as eap_chain.jl:117 (underflows n>~1100) */
  int32_t _pad;
  /* ---- clustering driver, mcmc_clustering_eap_chain.jl:19-153 ---- */
  double kappa, psi0;      /* --bend-mod, --bend-angle (eap_chain.jl:54-58,91-92)                    */
  double cutoff_radius;    /* --cutoff-radius in monomer lengths (UCutoff, eap_chain.jl:102,165-192) */
  double cluster_prob;     /* --cluster-prob: cluster_flip! returns early iff rand() <= it (:273)    */
  int32_t clustering;      /* 1: the trial of mcmc_clustering_eap_chain.jl:267-279 (move! + cluster_flip!,
                              alpha in the acceptor, two extra averagers)                           */
  int32_t alpha_carry;     /* 1 (reference): the acceptor stores logπ+log(alpha) as logπ_prev
                              (acceptance.jl:30-33); 0: stores logπ only (plain Metropolis-Hastings)  */
  int32_t cutoff_full;     /* 0 (reference): the UCutoff functor is the bare pair sum — no Σu, no −r·F
                              (eap_chain.jl:171-192 vs energy.jl:13-16); 1: Σu + U_cut − r·F          */
  int32_t planar;          /* 1: the 2-D tree (2D/inc/eap_chain.jl, 2D/mcmc_clustering_eap_chain.jl): phi-only state,
                              n = (cos phi, sin phi) in the x-z plane, no solid angle, flip_n! = phi + pi, the cluster
                              gate flips WITH probability cluster_prob, every stage starts from a new random chain   */
} orc_case;

typedef struct orc_chain orc_chain;

/* Philox4x32-10 (Salmon et al., SC'11), the counter RNG shared by oracle and CUDA path. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Stream definition shared with the CUDA path (DESIGN.md "RNG streams"). */
void orc_draw_init(uint64_t seed, uint32_t chain_id, uint32_t init, int64_t k, double* phi, double* theta);
void orc_draw_step(uint64_t seed, uint32_t chain_id, uint32_t init, int64_t step, int64_t n,
                   int64_t* idx0, double* u_phi, int32_t* flipbit, double* u_theta, double* eps);
double orc_draw_reinit_eps(uint64_t seed, uint32_t chain_id, uint32_t init);

/* Chain life cycle (inc/eap_chain.jl:60-135 ctor, :137-163 copy). */
orc_chain* orc_chain_new(const orc_case* c, const double* phi, const double* theta);
orc_chain* orc_chain_new_random(const orc_case* c, uint64_t seed, uint32_t chain_id, uint32_t init);
orc_chain* orc_chain_copy(const orc_chain* src);
void orc_chain_free(orc_chain* ch);

/* Accessors. out4 = {U, sum(us), U_dd (pair part), Omega}. */
void orc_chain_energy(const orc_chain* ch, double out4[4]);
double orc_chain_abs_pair_sum(const orc_chain* ch); /* sum |pair terms| for tolerance normalisation */
void orc_chain_r(const orc_chain* ch, double r[3]);          /* end_to_end, eap_chain.jl:405 */
void orc_chain_p(const orc_chain* ch, double p[3]);          /* chain_mu,  eap_chain.jl:408 */
void orc_chain_state(const orc_chain* ch, double* phi, double* theta);
void orc_chain_xs(const orc_chain* ch, double* xs /* 3n, column-major like the reference */);
void orc_chain_mus(const orc_chain* ch, double* mus /* 3n */);

/* move! (inc/eap_chain.jl:230-257): mutates, full energy recompute, idx0 is 0-based. */
void orc_chain_move(orc_chain* ch, int64_t idx0, double dphi, double dtheta);

/* Changed-pair ΔU (SURVEY §8a "ΔU decomposition"), non-mutating.
 * out = {dU, dOmega, sum|changed pair terms| (old and new), du_self, dr·F part, dU_pairs}. */
void orc_chain_delta_u(const orc_chain* ch, int64_t idx0, double dphi, double dtheta, double out[6]);

/* out8 = {U, sum(us) incl. bending, U_dd, Omega, U_bend, sum(psi)/(n-1), sum(cos^2 theta), sum|pair terms|}. */
void orc_chain_energy_ex(const orc_chain* ch, double out8[8]);
/* EAPChain(pargs) with --x0/--dx0 (eap_chain.jl:63-78): x0 has 2 or 2n entries (phi,theta interleaved);
 * the perturbations rand(Uniform(0,dx0[k])) reuse the uniforms of the random-init stream. */
orc_chain* orc_chain_new_x0(const orc_case* c, uint64_t seed, uint32_t chain_id, uint32_t init,
                            const double* x0, int64_t x0_len, const double dx0[2]);
/* Uniform #k of the cluster-growth streams (stream 0 = up, 1 = down) and the gate uniform of
 * cluster_flip! (eap_chain.jl:280,291,307). */
double orc_draw_cluster(uint64_t seed, uint32_t chain_id, uint32_t init, int64_t step, int32_t stream, int64_t k);
double orc_draw_cluster_gate(uint64_t seed, uint32_t chain_id, uint32_t init, int64_t step);
/* The composite trial of the clustering driver in changed-term form, non-mutating: move!(idx0,dphi,dtheta)
 * followed (if reflect) by refl_n! of every monomer in [lo0,hi0] (lo0 <= idx0 <= hi0).
 * out[12] = {dU, dOmega, sum|changed pair terms|, du_self, drF, dU_pairs, dU_bend, d sum(psi),
 *            d sum(cos^2 theta), dp1, dp2, dp3}. */
void orc_chain_delta_segment(const orc_chain* ch, int64_t idx0, double dphi, double dtheta, int32_t reflect,
                             int64_t lo0, int64_t hi0, double out[12]);
/* The same composite trial done literally as the reference does (move! then refl_n! per monomer, each a
 * full recompute); mutates. */
void orc_chain_move_segment(orc_chain* ch, int64_t idx0, double dphi, double dtheta, int32_t reflect,
                            int64_t lo0, int64_t hi0);
/* (1 + n̂_i·n̂_{i+1})/2, pflip_linear (eap_chain.jl:267,290). */
double orc_chain_link_prob(const orc_chain* ch, int64_t i0);

/* Whole MCMC loop, mcmc_eap_chain.jl:266-363 for ONE init segment of an existing run state. */
typedef struct orc_run orc_run;
orc_run* orc_run_new(const orc_case* c, uint64_t seed, uint32_t chain_id, int32_t algo /*0 full recompute, 1 ΔU*/);
void orc_run_set_state(orc_run* r, const double* phi, const double* theta);
/* Runs nsteps trials of the current init; rows may be NULL. traj: [nsteps/stepout][8], roll: [..][17]. */
void orc_run_steps(orc_run* r, int64_t nsteps, int64_t stepout, double* traj, double* roll);
/* Re-initialisation between inits (mcmc_eap_chain.jl:352-361); returns 1 if the new chain replaced the old. */
int32_t orc_run_reinit(orc_run* r, int32_t force_init);
/* Tape mode (tests only): the same Markov chain code fed from a caller-owned array of uniforms in the order the
 * reference's code calls rand() — initial chain, then per trial idx, dϕ, [Bool], dθ, [cluster_flip! draws], ϵ, and the
 * re-init draws — so that the oracle can be compared with the reference's own scripts run on a scripted `rand`
 * (tests/golden/ref_driver_*.jl).  reinit_stale = 1 keeps the acceptor's logπ_prev after a re-init swap, as the
 * reference does (mcmc_eap_chain.jl:360 vs inc/acceptance.jl:33); algo 0 only.  The tape must outlive the run.
 * orc_run_tape_pos: uniforms consumed so far, or -1 if the tape ran out. */
orc_run* orc_run_new_tape(const orc_case* c, int32_t algo, const double* tape, int64_t tape_len, int32_t reinit_stale);
int64_t orc_run_tape_pos(const orc_run* r);
/* avg[16] in rolling.csv column order (r1..r3, r1sq..r3sq, rsq, p1..p3, p1sq..p3sq, psq, U, Usq). */
void orc_run_averages(const orc_run* r, double avg[16], double* acc_rate, double* normalizer);
void orc_run_diag(const orc_run* r, double out[8]); /* phi_step, theta_step, nacc, natt, nacc_total, steps_total, U, Omega */
const orc_chain* orc_run_chain(const orc_run* r);
/* Clustering driver: a fresh `mcmc(nsteps, pargs, chain)` call on the current chain at temperature kT
 * (mcmc_clustering_eap_chain.jl:171-265: kT, U, weight function, acceptor, averagers, counters and step
 * sizes are rebuilt; the chain is kept).  Advances the RNG stream tag. */
void orc_run_begin_stage(orc_run* r, double kT);
/* Replace the chain by one built from --x0/--dx0. */
void orc_run_init_x0(orc_run* r, const double* x0, int64_t x0_len, const double dx0[2]);
/* The two extra averagers of the clustering driver (:243-244): <sum cos^2 theta>, <sum(psi)/(n-1)>. */
void orc_run_extra_averages(const orc_run* r, double ex[2]);
/* Clustering rows: roll19 = step + 16 + {Ealign, psi} (:334-346); state = [rows][2n] (phi,theta interleaved, :317). */
void orc_run_steps_ex(orc_run* r, int64_t nsteps, int64_t stepout, double* traj, double* roll19, double* state);
/* Cluster statistics of the run so far: {trials with a cluster flip, sum of cluster sizes, largest}. */
void orc_run_cluster_stats(const orc_run* r, double out[3]);
void orc_run_free(orc_run* r);

/* Multi-threaded throughput probe for bench.py's cpu_baseline / --impl reference leg:
 * nchains independent chains, one per thread at a time, nthreads pthreads; returns seconds. */
double orc_bench(const orc_case* c, uint64_t seed, int32_t algo, int32_t nchains, int64_t nsteps, int32_t nthreads);

#ifdef __cplusplus
}
#endif
#endif
