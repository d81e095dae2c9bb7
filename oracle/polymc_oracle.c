/*
 * polymc_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See polymc_oracle.h.
 *
 * PARITY UNPINNED by the reference (no tests / golden data upstream, Julia not installed here);
 * pinned by tests/golden/kat_n5.json (independent numpy restatement) and the
 * closed forms in oracle/closed_form.py (P2).
 *
 * All citations are into /root/reference.  Arithmetic is IEEE double throughout, as in the
 * reference.  Arrays 3×n are stored column-major ([3*i + k]) like Julia's.
 */
#include "polymc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3".   */
/* The reference draws from Julia's unseeded global RNG (mcmc_eap_chain.jl:277-287), so its     */
/* stream is unpinned by construction; the oracle and the CUDA path share this counter RNG.     */
/* ------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u53(uint32_t lo, uint32_t hi) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 11) * 0x1.0p-53; /* uniform on [0,1), 53 bits */
}

/* Stream tags (word 3 of the counter): (init << 8) | sub.                                      */
enum { SUB_STEP_A = 0, SUB_STEP_B = 1, SUB_INIT = 2, SUB_REINIT = 3, SUB_CLUSTER_UP = 4, SUB_CLUSTER_DOWN = 5,
       SUB_CLUSTER_GATE = 6 };

static void philox_at(uint64_t seed, uint32_t chain_id, uint32_t init, uint32_t sub, uint64_t pos,
                      uint32_t out[4]) {
  uint32_t ctr[4] = {(uint32_t)pos, (uint32_t)(pos >> 32), chain_id, (init << 8) | sub};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  orc_philox4x32_10(ctr, key, out);
}

/* Random initial angles: phi ~ U(0,2pi), theta ~ U(0,pi) (inc/eap_chain.jl:6-7,62). */
void orc_draw_init(uint64_t seed, uint32_t chain_id, uint32_t init, int64_t k, double* phi, double* theta) {
  uint32_t w[4];
  philox_at(seed, chain_id, init, SUB_INIT, (uint64_t)k, w);
  *phi = 0.0 + (2 * M_PI - 0.0) * u53(w[0], w[1]);
  *theta = 0.0 + (M_PI - 0.0) * u53(w[2], w[3]);
}

/* Draw order of one trial: idx, dphi, [Bool], dtheta, eps (mcmc_eap_chain.jl:277-280,287). */
void orc_draw_step(uint64_t seed, uint32_t chain_id, uint32_t init, int64_t step, int64_t n, int64_t* idx0,
                   double* u_phi, int32_t* flipbit, double* u_theta, double* eps) {
  uint32_t a[4], b[4];
  philox_at(seed, chain_id, init, SUB_STEP_A, (uint64_t)step, a);
  philox_at(seed, chain_id, init, SUB_STEP_B, (uint64_t)step, b);
  uint64_t v = ((uint64_t)a[1] << 32) | a[0];
  *idx0 = (int64_t)(((unsigned __int128)v * (unsigned __int128)(uint64_t)n) >> 64);
  *u_phi = u53(a[2], a[3]);
  *flipbit = (int32_t)(a[2] & 1u);
  *u_theta = u53(b[0], b[1]);
  *eps = u53(b[2], b[3]);
}

/* cluster_flip! draws an unbounded number of uniforms per trial (eap_chain.jl:273,291,307).  Uniform #k
 * of the upward (stream 0) / downward (stream 1) growth is word pair (k&1) of the Philox block at
 * position step + ((k>>1) << 40); steps stay below 2^40. */
double orc_draw_cluster(uint64_t seed, uint32_t chain_id, uint32_t init, int64_t step, int32_t stream, int64_t k) {
  uint32_t w[4];
  philox_at(seed, chain_id, init, stream ? SUB_CLUSTER_DOWN : SUB_CLUSTER_UP,
            (uint64_t)step + (((uint64_t)k >> 1) << 40), w);
  return (k & 1) ? u53(w[2], w[3]) : u53(w[0], w[1]);
}

double orc_draw_cluster_gate(uint64_t seed, uint32_t chain_id, uint32_t init, int64_t step) {
  uint32_t w[4];
  philox_at(seed, chain_id, init, SUB_CLUSTER_GATE, (uint64_t)step, w);
  return u53(w[0], w[1]);
}

double orc_draw_reinit_eps(uint64_t seed, uint32_t chain_id, uint32_t init) {
  uint32_t w[4];
  philox_at(seed, chain_id, init, SUB_REINIT, 0, w);
  return u53(w[0], w[1]);
}

/* ------------------------------------------------------------------------------------------ */
/* EAPChain (inc/eap_chain.jl:12-36).  kappa == 0 in mcmc_eap_chain.jl (no --bend-mod; :91); the */
/* clustering driver sets it, and averages psi (mcmc_clustering_eap_chain.jl:244).              */
/* ------------------------------------------------------------------------------------------ */
struct orc_chain {
  orc_case c;
  int64_t n;
  double *phi, *cphi, *sphi, *theta, *ctheta, *stheta;
  double *nhat, *mus, *us, *xs; /* 3n, 3n, n, 3n */
  double *psis;                 /* n-1 bond angles (eap_chain.jl:29) */
  double r[3];
  double Omega;
  double U;
};

/* n̂ (eap_chain.jl:40) */
static inline void nhat_of(double cphi, double sphi, double cth, double sth, double out[3]) {
  out[0] = cphi * sth;
  out[1] = sphi * sth;
  out[2] = cth;
}

/* Dipole response (inc/dipole_response.jl:7-11,18-21 dielectric; :27-29 polar with M = mu*I,
 * eap_chain.jl:84). */
static inline void mu_of(const orc_case* c, double cphi, double sphi, double cth, double sth, double out[3]) {
  double nh[3];
  nhat_of(cphi, sphi, cth, sth, nh);
  if (c->chain_type == ORC_CHAIN_DIELECTRIC) {
    double f = (c->K1 - c->K2) * c->E0 * cth;
    out[0] = f * nh[0] + c->K2 * 0.0;
    out[1] = f * nh[1] + c->K2 * 0.0;
    out[2] = f * nh[2] + c->K2 * c->E0;
  } else {
    out[0] = c->mu * nh[0];
    out[1] = c->mu * nh[1];
    out[2] = c->mu * nh[2];
  }
}

/* The 2-D tree: n̂ = (cosϕ, sinϕ) (2D/inc/eap_chain.jl:33) with the field along the SECOND axis
 * (μ_dielectric = (K1−K2)·E0·sinϕ·n̂ + K2·[0;E0], 2D/inc/dipole_response.jl:7-10; polar M·n̂, :25-27;
 * u = −½E0μ[2], 2D/inc/eap_chain.jl:64).  Stored here in the x–z plane of the 3-vectors, y ≡ 0. */
static inline void nhat_planar(double cphi, double sphi, double out[3]) {
  out[0] = cphi;
  out[1] = 0.0;
  out[2] = sphi;
}
static inline void mu_planar(const orc_case* c, double cphi, double sphi, double out[3]) {
  double nh[3];
  nhat_planar(cphi, sphi, nh);
  if (c->chain_type == ORC_CHAIN_DIELECTRIC) {
    double f = (c->K1 - c->K2) * c->E0 * sphi;
    out[0] = f * nh[0] + c->K2 * 0.0;
    out[1] = 0.0;
    out[2] = f * nh[2] + c->K2 * c->E0;
  } else {
    out[0] = c->mu * nh[0];
    out[1] = 0.0;
    out[2] = c->mu * nh[2];
  }
}
/* direction and dipole of one monomer in either tree */
static inline void dir_c(const orc_case* c, double cphi, double sphi, double cth, double sth, double nh[3], double mu[3]) {
  if (c->planar) {
    nhat_planar(cphi, sphi, nh);
    mu_planar(c, cphi, sphi, mu);
  } else {
    nhat_of(cphi, sphi, cth, sth, nh);
    mu_of(c, cphi, sphi, cth, sth, mu);
  }
}

/* u(E0, mu) = -1/2 E0 mu_z  (eap_chain.jl:53); same 1/2 for polar chains. */
static inline double u_self(double E0, const double mu[3]) { return -1.0 / 2 * E0 * mu[2]; }

/* ψj (eap_chain.jl:45-47): acos(min(1, max(-1, n̂_i·n̂_{i+1}))) */
static inline double psi_of(const double* na, const double* nb) {
  double d = na[0] * nb[0] + na[1] * nb[1] + na[2] * nb[2];
  return acos(fmin(1.0, fmax(-1.0, d)));
}
static inline double psi_j(const orc_chain* ch, int64_t i) { return psi_of(&ch->nhat[3 * i], &ch->nhat[3 * (i + 1)]); }

/* ubend (eap_chain.jl:54-58): κ/2 (ψ_i − ψ0)² for every monomer but the last. */
static inline double ubend_val(const orc_case* c, double psi) { return c->kappa / 2 * (psi - c->psi0) * (psi - c->psi0); }
static inline double ubend(const orc_chain* ch, int64_t i) {
  return (i != ch->n - 1) ? ubend_val(&ch->c, ch->psis[i]) : 0.0;
}

/* update_xs! (eap_chain.jl:49-51): xs = b (cumsum(n̂) - n̂/2) */
static void update_xs(orc_chain* ch) {
  double s[3] = {0, 0, 0};
  for (int64_t i = 0; i < ch->n; ++i)
    for (int k = 0; k < 3; ++k) {
      s[k] += ch->nhat[3 * i + k];
      ch->xs[3 * i + k] = ch->c.b * (s[k] - 0.5 * ch->nhat[3 * i + k]);
    }
}

/* end_to_end (eap_chain.jl:405-406): xs[:,end] + b/2 n̂_n */
static void end_to_end(const orc_chain* ch, double r[3]) {
  int64_t l = ch->n - 1;
  for (int k = 0; k < 3; ++k) r[k] = ch->xs[3 * l + k] + ch->c.b / 2.0 * ch->nhat[3 * l + k];
}

/* One dipole-dipole pair term (eap_chain.jl:200-207). */
static inline double pair_term(const double* xi, const double* xj, const double* mi, const double* mj) {
  double r0 = xi[0] - xj[0], r1 = xi[1] - xj[1], r2c = xi[2] - xj[2];
  double r2 = r0 * r0 + r1 * r1 + r2c * r2c;
  double rmag = sqrt(r2);
  double h0 = r0 / rmag, h1 = r1 / rmag, h2 = r2c / rmag;
  double r3 = r2 * rmag;
  double mm = mi[0] * mj[0] + mi[1] * mj[1] + mi[2] * mj[2];
  double a = mi[0] * h0 + mi[1] * h1 + mi[2] * h2;
  double b = mj[0] * h0 + mj[1] * h1 + mj[2] * h2;
  return (mm - 3 * a * b) / (4 * M_PI * r3);
}

/* One term of UCutoff (eap_chain.jl:176-187): zero beyond the cut-off radius. */
static inline double pair_term_cut(const double* xi, const double* xj, const double* mi, const double* mj,
                                   double crad2) {
  double r0 = xi[0] - xj[0], r1 = xi[1] - xj[1], r2c = xi[2] - xj[2];
  double r2 = r0 * r0 + r1 * r1 + r2c * r2c;
  if (r2 > crad2) return 0.0;
  return pair_term(xi, xj, mi, mj);
}

static inline double crad2_of(const orc_case* c) {
  double cr = c->cutoff_radius * c->b; /* UCutoff(pargs["cutoff-radius"]*pargs["mlen"]), eap_chain.jl:102 */
  return cr * cr;
}

/* The pair term the chain's energy type uses between two arbitrary sites. */
static inline double pair_any(const orc_case* c, const double* xi, const double* xj, const double* mi,
                              const double* mj) {
  if (c->energy_type == ORC_ENERGY_CUTOFF) return pair_term_cut(xi, xj, mi, mj, crad2_of(c));
  return pair_term(xi, xj, mi, mj);
}

/* UCutoff functor (eap_chain.jl:171-192). */
static double U_cutoff(const orc_chain* ch) {
  double crad2 = crad2_of(&ch->c);
  double U = 0.0;
  for (int64_t i = 0; i < ch->n; ++i)
    for (int64_t j = i + 1; j < ch->n; ++j)
      U += pair_term_cut(&ch->xs[3 * i], &ch->xs[3 * j], &ch->mus[3 * i], &ch->mus[3 * j], crad2);
  return U;
}

/* U_interaction (eap_chain.jl:196-211): all pairs, i outer, j inner. */
static double U_interaction(const orc_chain* ch) {
  double U = 0.0;
  for (int64_t i = 0; i < ch->n; ++i)
    for (int64_t j = i + 1; j < ch->n; ++j)
      U += pair_term(&ch->xs[3 * i], &ch->xs[3 * j], &ch->mus[3 * i], &ch->mus[3 * j]);
  return U;
}

/* U_Ising (eap_chain.jl:215-228): nearest neighbours only. */
static double U_Ising(const orc_chain* ch) {
  double U = 0.0;
  for (int64_t i = 0; i + 1 < ch->n; ++i)
    U += pair_term(&ch->xs[3 * i], &ch->xs[3 * (i + 1)], &ch->mus[3 * i], &ch->mus[3 * (i + 1)]);
  return U;
}

static double sum_us(const orc_chain* ch) {
  double s = 0.0;
  for (int64_t i = 0; i < ch->n; ++i) s += ch->us[i];
  return s;
}

static double U_pairs(const orc_chain* ch) {
  if (ch->c.energy_type == ORC_ENERGY_INTERACTING) return U_interaction(ch);
  if (ch->c.energy_type == ORC_ENERGY_ISING) return U_Ising(ch);
  if (ch->c.energy_type == ORC_ENERGY_CUTOFF) return U_cutoff(ch);
  return 0.0;
}

/* Energy functors (inc/energy.jl:7-9, :13-16, :20-23) and UCutoff (eap_chain.jl:171-192), which is the
 * bare pair sum: it adds neither Σu nor −r·F (unlike InteractingEnergy). */
static double U_total(const orc_chain* ch) {
  double r[3];
  end_to_end(ch, r);
  double rf = r[0] * ch->c.Fx + r[1] * 0.0 + r[2] * ch->c.Fz;
  if (ch->c.energy_type == ORC_ENERGY_NONINTERACTING) return sum_us(ch) - rf;
  if (ch->c.energy_type == ORC_ENERGY_CUTOFF && !ch->c.cutoff_full) return U_cutoff(ch);
  return sum_us(ch) + U_pairs(ch) - rf;
}

static orc_chain* chain_alloc(const orc_case* c) {
  orc_chain* ch = (orc_chain*)calloc(1, sizeof(orc_chain));
  ch->c = *c;
  ch->n = c->n;
  size_t n = (size_t)c->n;
  ch->phi = (double*)calloc(n, sizeof(double));
  ch->cphi = (double*)calloc(n, sizeof(double));
  ch->sphi = (double*)calloc(n, sizeof(double));
  ch->theta = (double*)calloc(n, sizeof(double));
  ch->ctheta = (double*)calloc(n, sizeof(double));
  ch->stheta = (double*)calloc(n, sizeof(double));
  ch->nhat = (double*)calloc(3 * n, sizeof(double));
  ch->mus = (double*)calloc(3 * n, sizeof(double));
  ch->us = (double*)calloc(n, sizeof(double));
  ch->xs = (double*)calloc(3 * n, sizeof(double));
  ch->psis = (double*)calloc(n ? n : 1, sizeof(double));
  return ch;
}

/* EAPChain(pargs) (eap_chain.jl:60-135) with the angles supplied by the caller. */
orc_chain* orc_chain_new(const orc_case* c, const double* phi, const double* theta) {
  orc_chain* ch = chain_alloc(c);
  double prod = 1.0, slog = 0.0;
  for (int64_t i = 0; i < ch->n; ++i) {
    ch->phi[i] = phi[i];
    ch->theta[i] = theta[i];
    ch->cphi[i] = cos(phi[i]);
    ch->sphi[i] = sin(phi[i]);
    ch->ctheta[i] = cos(theta[i]);
    ch->stheta[i] = sin(theta[i]);
    if (c->planar) { /* no θ in the 2-D tree: sinθ ≡ 1 makes every solid-angle term exactly 0 */
      ch->theta[i] = 0.0;
      ch->ctheta[i] = ch->sphi[i];
      ch->stheta[i] = 1.0;
    }
    prod *= ch->stheta[i];
    slog += log(ch->stheta[i]);
    dir_c(c, ch->cphi[i], ch->sphi[i], ch->ctheta[i], ch->stheta[i], &ch->nhat[3 * i], &ch->mus[3 * i]);
  }
  for (int64_t i = 0; i + 1 < ch->n; ++i) ch->psis[i] = psi_j(ch, i);                           /* :125 */
  for (int64_t i = 0; i < ch->n; ++i) ch->us[i] = u_self(c->E0, &ch->mus[3 * i]) + ubend(ch, i); /* :130 */
  /* eap_chain.jl:117 stores log(prod(sin θ)), which underflows to -Inf for n >~ 1100 and then
   * accepts every move (SURVEY finding 7).  The intended Σ log sin θ is the default here.     */
  ch->Omega = c->omega_compat ? log(prod) : slog;
  update_xs(ch);
  end_to_end(ch, ch->r);
  ch->U = U_total(ch);
  return ch;
}

orc_chain* orc_chain_new_random(const orc_case* c, uint64_t seed, uint32_t chain_id, uint32_t init) {
  double* phi = (double*)malloc(sizeof(double) * (size_t)c->n);
  double* theta = (double*)malloc(sizeof(double) * (size_t)c->n);
  for (int64_t k = 0; k < c->n; ++k) orc_draw_init(seed, chain_id, init, k, &phi[k], &theta[k]);
  orc_chain* ch = orc_chain_new(c, phi, theta);
  free(phi);
  free(theta);
  return ch;
}

/* EAPChain(pargs) with --x0/--dx0 (eap_chain.jl:63-78): ϕ = ϕ0 + rand(Uniform(0,dx0[1])), θ = θ0 +
 * rand(Uniform(0,dx0[2])); x0 = [ϕ;θ] for all monomers or 2n interleaved values. */
orc_chain* orc_chain_new_x0(const orc_case* c, uint64_t seed, uint32_t chain_id, uint32_t init, const double* x0,
                            int64_t x0_len, const double dx0[2]) {
  if (x0_len != 2 && x0_len != 2 * c->n) return NULL;
  double* phi = (double*)malloc(sizeof(double) * (size_t)c->n);
  double* theta = (double*)malloc(sizeof(double) * (size_t)c->n);
  for (int64_t k = 0; k < c->n; ++k) {
    uint32_t w[4];
    philox_at(seed, chain_id, init, SUB_INIT, (uint64_t)k, w);
    double p0 = x0_len == 2 ? x0[0] : x0[2 * k], t0 = x0_len == 2 ? x0[1] : x0[2 * k + 1];
    phi[k] = p0 + (0.0 + (dx0[0] - 0.0) * u53(w[0], w[1]));
    theta[k] = t0 + (0.0 + (dx0[1] - 0.0) * u53(w[2], w[3]));
  }
  orc_chain* ch = orc_chain_new(c, phi, theta);
  free(phi);
  free(theta);
  return ch;
}

/* EAPChain(chain) deep copy (eap_chain.jl:137-163). */
orc_chain* orc_chain_copy(const orc_chain* s) {
  orc_chain* ch = chain_alloc(&s->c);
  size_t n = (size_t)s->n;
  memcpy(ch->phi, s->phi, n * sizeof(double));
  memcpy(ch->cphi, s->cphi, n * sizeof(double));
  memcpy(ch->sphi, s->sphi, n * sizeof(double));
  memcpy(ch->theta, s->theta, n * sizeof(double));
  memcpy(ch->ctheta, s->ctheta, n * sizeof(double));
  memcpy(ch->stheta, s->stheta, n * sizeof(double));
  memcpy(ch->nhat, s->nhat, 3 * n * sizeof(double));
  memcpy(ch->mus, s->mus, 3 * n * sizeof(double));
  memcpy(ch->us, s->us, n * sizeof(double));
  memcpy(ch->xs, s->xs, 3 * n * sizeof(double));
  memcpy(ch->psis, s->psis, n * sizeof(double));
  memcpy(ch->r, s->r, sizeof(ch->r));
  ch->Omega = s->Omega;
  ch->U = s->U;
  return ch;
}

static void chain_assign(orc_chain* d, const orc_chain* s) {
  size_t n = (size_t)s->n;
  memcpy(d->phi, s->phi, n * sizeof(double));
  memcpy(d->cphi, s->cphi, n * sizeof(double));
  memcpy(d->sphi, s->sphi, n * sizeof(double));
  memcpy(d->theta, s->theta, n * sizeof(double));
  memcpy(d->ctheta, s->ctheta, n * sizeof(double));
  memcpy(d->stheta, s->stheta, n * sizeof(double));
  memcpy(d->nhat, s->nhat, 3 * n * sizeof(double));
  memcpy(d->mus, s->mus, 3 * n * sizeof(double));
  memcpy(d->us, s->us, n * sizeof(double));
  memcpy(d->xs, s->xs, 3 * n * sizeof(double));
  memcpy(d->psis, s->psis, n * sizeof(double));
  memcpy(d->r, s->r, sizeof(d->r));
  d->Omega = s->Omega;
  d->U = s->U;
}

void orc_chain_free(orc_chain* ch) {
  if (!ch) return;
  free(ch->phi); free(ch->cphi); free(ch->sphi);
  free(ch->theta); free(ch->ctheta); free(ch->stheta);
  free(ch->nhat); free(ch->mus); free(ch->us); free(ch->xs); free(ch->psis);
  free(ch);
}

void orc_chain_energy(const orc_chain* ch, double out4[4]) {
  out4[0] = U_total(ch);
  out4[1] = sum_us(ch);
  out4[2] = U_pairs(ch);
  out4[3] = ch->Omega;
}

double orc_chain_abs_pair_sum(const orc_chain* ch) {
  double s = 0.0;
  if (ch->c.energy_type == ORC_ENERGY_INTERACTING || ch->c.energy_type == ORC_ENERGY_CUTOFF) {
    for (int64_t i = 0; i < ch->n; ++i)
      for (int64_t j = i + 1; j < ch->n; ++j)
        s += fabs(pair_any(&ch->c, &ch->xs[3 * i], &ch->xs[3 * j], &ch->mus[3 * i], &ch->mus[3 * j]));
  } else if (ch->c.energy_type == ORC_ENERGY_ISING) {
    for (int64_t i = 0; i + 1 < ch->n; ++i)
      s += fabs(pair_term(&ch->xs[3 * i], &ch->xs[3 * (i + 1)], &ch->mus[3 * i], &ch->mus[3 * (i + 1)]));
  }
  return s;
}

static double sum_ubend(const orc_chain* ch) {
  double s = 0.0;
  for (int64_t i = 0; i + 1 < ch->n; ++i) s += ubend(ch, i);
  return s;
}

/* accessors of the two extra averagers, mcmc_clustering_eap_chain.jl:243-244 */
static double sum_cos2(const orc_chain* ch) {
  double s = 0.0;
  for (int64_t i = 0; i < ch->n; ++i) s += ch->ctheta[i] * ch->ctheta[i];
  return s;
}
static double sum_psi(const orc_chain* ch) {
  double s = 0.0;
  for (int64_t i = 0; i + 1 < ch->n; ++i) s += ch->psis[i];
  return s;
}

void orc_chain_energy_ex(const orc_chain* ch, double out8[8]) {
  out8[0] = U_total(ch);
  out8[1] = sum_us(ch);
  out8[2] = U_pairs(ch);
  out8[3] = ch->Omega;
  out8[4] = sum_ubend(ch);
  out8[5] = sum_psi(ch) / (double)(ch->n - 1);
  out8[6] = sum_cos2(ch);
  out8[7] = orc_chain_abs_pair_sum(ch);
}

double orc_chain_link_prob(const orc_chain* ch, int64_t i) {
  const double *a = &ch->nhat[3 * i], *b = &ch->nhat[3 * (i + 1)];
  return (1 + (a[0] * b[0] + a[1] * b[1] + a[2] * b[2])) / 2;
}

void orc_chain_r(const orc_chain* ch, double r[3]) { end_to_end(ch, r); }

/* chain_μ (eap_chain.jl:408): Σ_i μ_i */
void orc_chain_p(const orc_chain* ch, double p[3]) {
  p[0] = p[1] = p[2] = 0.0;
  for (int64_t i = 0; i < ch->n; ++i)
    for (int k = 0; k < 3; ++k) p[k] += ch->mus[3 * i + k];
}

void orc_chain_state(const orc_chain* ch, double* phi, double* theta) {
  memcpy(phi, ch->phi, sizeof(double) * (size_t)ch->n);
  memcpy(theta, ch->theta, sizeof(double) * (size_t)ch->n);
}
void orc_chain_xs(const orc_chain* ch, double* xs) { memcpy(xs, ch->xs, sizeof(double) * 3 * (size_t)ch->n); }
void orc_chain_mus(const orc_chain* ch, double* mus) { memcpy(mus, ch->mus, sizeof(double) * 3 * (size_t)ch->n); }

/* move! (eap_chain.jl:230-257) up to and including `chain.r[:] = end_to_end(chain)` (:253).
 * ϕ is not wrapped (:232); θ is clamped, not reflected (:236). */
static void chain_move_caches(orc_chain* ch, int64_t idx, double dphi, double dtheta) {
  ch->phi[idx] += dphi;
  ch->cphi[idx] = cos(ch->phi[idx]);
  ch->sphi[idx] = sin(ch->phi[idx]);
  if (ch->c.planar) { /* move!(chain, idx, dϕ), 2D/inc/eap_chain.jl:171-187 */
    ch->ctheta[idx] = ch->sphi[idx];
  } else {
    ch->theta[idx] = fmin(M_PI, fmax(0.0, ch->theta[idx] + dtheta));
    double sth = sin(ch->theta[idx]);
    ch->Omega += log(sth / ch->stheta[idx]);
    ch->ctheta[idx] = cos(ch->theta[idx]);
    ch->stheta[idx] = sth;
  }
  dir_c(&ch->c, ch->cphi[idx], ch->sphi[idx], ch->ctheta[idx], ch->stheta[idx], &ch->nhat[3 * idx], &ch->mus[3 * idx]);
  if (idx < ch->n - 1) ch->psis[idx] = psi_j(ch, idx);                                      /* :246 */
  if (idx > 0) {                                                                            /* :247-250 */
    ch->psis[idx - 1] = psi_j(ch, idx - 1);
    ch->us[idx - 1] = u_self(ch->c.E0, &ch->mus[3 * (idx - 1)]) + ubend(ch, idx - 1);
  }
  ch->us[idx] = u_self(ch->c.E0, &ch->mus[3 * idx]) + ubend(ch, idx);                       /* :251 */
  update_xs(ch);
  end_to_end(ch, ch->r);
}

/* move! (eap_chain.jl:230-257): caches, then the full energy recompute of :254. */
void orc_chain_move(orc_chain* ch, int64_t idx, double dphi, double dtheta) {
  chain_move_caches(ch, idx, dphi, dtheta);
  ch->U = U_total(ch);
}

/* Changed-pair ΔU: equals U(move!(copy)) - U(chain) in real arithmetic. */
void orc_chain_delta_u(const orc_chain* ch, int64_t idx, double dphi, double dtheta, double out[6]) {
  const orc_case* c = &ch->c;
  int64_t n = ch->n;
  double phi1 = ch->phi[idx] + dphi;
  double th1 = fmin(M_PI, fmax(0.0, ch->theta[idx] + dtheta));
  double cph = cos(phi1), sph = sin(phi1), cth = cos(th1), sth = sin(th1);
  double nh1[3], mu1[3];
  nhat_of(cph, sph, cth, sth, nh1);
  mu_of(c, cph, sph, cth, sth, mu1);
  double dn[3], D[3], xi_new[3];
  for (int k = 0; k < 3; ++k) {
    dn[k] = nh1[k] - ch->nhat[3 * idx + k];
    D[k] = c->b * dn[k];                                   /* tail translation            */
    xi_new[k] = ch->xs[3 * idx + k] + 0.5 * c->b * dn[k];  /* x'_idx = x_idx + (b/2)Δn̂   */
  }
  double dOmega = log(sth / ch->stheta[idx]);
  double du = u_self(c->E0, mu1) - u_self(c->E0, &ch->mus[3 * idx]);
  double drF = -(D[0] * c->Fx + D[2] * c->Fz);
  double dpair = 0.0, abs_sum = 0.0;
  const double* xi_old = &ch->xs[3 * idx];
  const double* mi_old = &ch->mus[3 * idx];
  if (c->energy_type != ORC_ENERGY_NONINTERACTING) {
    int64_t jlo = 0, jhi = n - 1;
    if (c->energy_type == ORC_ENERGY_ISING) {
      jlo = idx > 0 ? idx - 1 : 0;
      jhi = idx + 1 < n ? idx + 1 : n - 1;
    }
    /* row: pairs (idx, j) */
    for (int64_t j = jlo; j <= jhi; ++j) {
      if (j == idx) continue;
      double xj_new[3];
      for (int k = 0; k < 3; ++k) xj_new[k] = ch->xs[3 * j + k] + (j > idx ? D[k] : 0.0);
      double e_old = pair_term(xi_old, &ch->xs[3 * j], mi_old, &ch->mus[3 * j]);
      double e_new = pair_term(xi_new, xj_new, mu1, &ch->mus[3 * j]);
      dpair += e_new - e_old;
      abs_sum += fabs(e_new) + fabs(e_old);
    }
    /* rectangle: heads i < idx against tails j > idx (all-pairs energy only) */
    if (c->energy_type == ORC_ENERGY_INTERACTING) {
      for (int64_t i = 0; i < idx; ++i)
        for (int64_t j = idx + 1; j < n; ++j) {
          double xj_new[3];
          for (int k = 0; k < 3; ++k) xj_new[k] = ch->xs[3 * j + k] + D[k];
          double e_old = pair_term(&ch->xs[3 * i], &ch->xs[3 * j], &ch->mus[3 * i], &ch->mus[3 * j]);
          double e_new = pair_term(&ch->xs[3 * i], xj_new, &ch->mus[3 * i], &ch->mus[3 * j]);
          dpair += e_new - e_old;
          abs_sum += fabs(e_new) + fabs(e_old);
        }
    }
  }
  out[0] = du + drF + dpair;
  out[1] = dOmega;
  out[2] = abs_sum;
  out[3] = du;
  out[4] = drF;
  out[5] = dpair;
}

/* refl_n! (eap_chain.jl:263-265): move!(chain, idx, 0, π − 2θ_idx). */
static void chain_refl_caches(orc_chain* ch, int64_t i) {
  if (ch->c.planar) chain_move_caches(ch, i, M_PI, 0.0); /* flip_n! = move!(chain, idx, π), 2D/inc/eap_chain.jl:189-191 */
  else chain_move_caches(ch, i, 0.0, M_PI - 2 * ch->theta[i]);
}

/* The composite trial of the clustering driver done literally (mcmc_clustering_eap_chain.jl:272-273 →
 * eap_chain.jl:230-257, :311-315): move!, then refl_n! for every monomer of the cluster, each with its
 * own full energy recompute. */
void orc_chain_move_segment(orc_chain* ch, int64_t idx, double dphi, double dtheta, int32_t reflect, int64_t lo,
                            int64_t hi) {
  orc_chain_move(ch, idx, dphi, dtheta);
  if (reflect)
    for (int64_t i = lo; i <= hi; ++i) {
      chain_refl_caches(ch, i);
      ch->U = U_total(ch);
    }
}

/* The same composite trial in changed-term form (ours; equal to the literal form in real arithmetic).
 * Monomers lo..hi get new directions and dipoles, the tail j > hi is translated rigidly by D = bΣΔn̂, so
 * the changed pair terms are segment×everything and heads(i<lo)×tails(j>hi); bond angles ψ change on the
 * bonds lo−1..hi. */
void orc_chain_delta_segment(const orc_chain* ch, int64_t idx, double dphi, double dtheta, int32_t reflect,
                             int64_t lo, int64_t hi, double out[12]) {
  const orc_case* c = &ch->c;
  const int64_t n = ch->n;
  if (!reflect) lo = hi = idx;
  const int64_t m = hi - lo + 1;
  double* nn = (double*)malloc(sizeof(double) * 3 * (size_t)m);  /* n̂' */
  double* mn = (double*)malloc(sizeof(double) * 3 * (size_t)m);  /* μ' */
  double* xn = (double*)malloc(sizeof(double) * 3 * (size_t)m);  /* x' */
  double dOmega = 0.0, du = 0.0, dcos2 = 0.0, dp[3] = {0, 0, 0};
  double S[3] = {0, 0, 0};  /* running Σ Δn̂ over the segment */
  for (int64_t k = 0; k < m; ++k) {
    const int64_t i = lo + k;
    double phi1 = ch->phi[i], th1 = ch->theta[i], sprev = ch->stheta[i];
    if (c->planar) { /* 2-D: move! adds dϕ, flip_n! adds π; no θ, no solid angle */
      if (i == idx) phi1 += dphi;
      if (reflect) phi1 += M_PI;
    } else {
      if (i == idx) {  /* move! (:232-238) */
        phi1 += dphi;
        th1 = fmin(M_PI, fmax(0.0, th1 + dtheta));
        double s1 = sin(th1);
        dOmega += log(s1 / sprev);
        sprev = s1;
      }
      if (reflect) {   /* refl_n!: dϕ = 0, dθ = π − 2θ (:263-265) */
        phi1 += 0.0;
        th1 = fmin(M_PI, fmax(0.0, th1 + (M_PI - 2 * th1)));
        double s2 = sin(th1);
        dOmega += log(s2 / sprev);
        sprev = s2;
      }
    }
    double cph = cos(phi1), sph = sin(phi1), cth = cos(th1), sth = sin(th1);
    if (c->planar) { cth = sph; sth = 1.0; }
    dir_c(c, cph, sph, cth, sth, &nn[3 * k], &mn[3 * k]);
    du += u_self(c->E0, &mn[3 * k]) - u_self(c->E0, &ch->mus[3 * i]);
    dcos2 += cth * cth - ch->ctheta[i] * ch->ctheta[i];
    for (int q = 0; q < 3; ++q) {
      const double dn = nn[3 * k + q] - ch->nhat[3 * i + q];
      xn[3 * k + q] = ch->xs[3 * i + q] + c->b * (S[q] + 0.5 * dn);  /* update_xs! restricted to the segment */
      S[q] += dn;
      dp[q] += mn[3 * k + q] - ch->mus[3 * i + q];
    }
  }
  const double D[3] = {c->b * S[0], c->b * S[1], c->b * S[2]};
  const double drF = -(D[0] * c->Fx + D[2] * c->Fz);
  /* bonds lo−1 .. hi */
  double dbend = 0.0, dpsi = 0.0;
  for (int64_t i = (lo > 0 ? lo - 1 : 0); i <= hi && i + 1 < n; ++i) {
    const double* a = (i >= lo) ? &nn[3 * (i - lo)] : &ch->nhat[3 * i];
    const double* b2 = (i + 1 <= hi) ? &nn[3 * (i + 1 - lo)] : &ch->nhat[3 * (i + 1)];
    const double psi_new = psi_of(a, b2);
    dpsi += psi_new - ch->psis[i];
    dbend += ubend_val(c, psi_new) - ubend_val(c, ch->psis[i]);
  }
  /* pair terms */
  double dpair = 0.0, abs_sum = 0.0;
#define NEW_X(j, buf)                                                                     \
  const double* buf##p;                                                                   \
  double buf[3];                                                                          \
  if ((j) >= lo && (j) <= hi) buf##p = &xn[3 * ((j) - lo)];                               \
  else if ((j) > hi) { for (int q = 0; q < 3; ++q) buf[q] = ch->xs[3 * (j) + q] + D[q]; buf##p = buf; } \
  else buf##p = &ch->xs[3 * (j)];
#define NEW_MU(j) (((j) >= lo && (j) <= hi) ? &mn[3 * ((j) - lo)] : &ch->mus[3 * (j)])
#define ADD_PAIR(i, j)                                                                     \
  do {                                                                                     \
    NEW_X(i, xi) NEW_X(j, xj)                                                              \
    const double e_old = pair_any(c, &ch->xs[3 * (i)], &ch->xs[3 * (j)], &ch->mus[3 * (i)], &ch->mus[3 * (j)]); \
    const double e_new = pair_any(c, xip, xjp, NEW_MU(i), NEW_MU(j));                      \
    dpair += e_new - e_old;                                                                \
    abs_sum += fabs(e_new) + fabs(e_old);                                                  \
  } while (0)
  if (c->energy_type == ORC_ENERGY_ISING) {
    for (int64_t i = (lo > 0 ? lo - 1 : 0); i <= hi && i + 1 < n; ++i) ADD_PAIR(i, i + 1);
  } else if (c->energy_type == ORC_ENERGY_INTERACTING || c->energy_type == ORC_ENERGY_CUTOFF) {
    for (int64_t s2 = lo; s2 <= hi; ++s2) {  /* segment × everything (each pair once) */
      for (int64_t j = 0; j < lo; ++j) ADD_PAIR(j, s2);
      for (int64_t j = s2 + 1; j < n; ++j) ADD_PAIR(s2, j);
    }
    for (int64_t i = 0; i < lo; ++i)         /* heads × tails */
      for (int64_t j = hi + 1; j < n; ++j) ADD_PAIR(i, j);
  }
#undef ADD_PAIR
#undef NEW_MU
#undef NEW_X
  const int bare = (c->energy_type == ORC_ENERGY_CUTOFF && !c->cutoff_full);
  out[0] = bare ? dpair : (du + dbend) + drF + dpair;
  out[1] = dOmega;
  out[2] = abs_sum;
  out[3] = du;
  out[4] = drF;
  out[5] = dpair;
  out[6] = dbend;
  out[7] = dpsi;
  out[8] = dcos2;
  out[9] = dp[0]; out[10] = dp[1]; out[11] = dp[2];
  free(nn); free(mn); free(xn);
}

/* ------------------------------------------------------------------------------------------ */
/* The sampler: Metropolis (inc/acceptance.jl:13-39), adaptation (mcmc_eap_chain.jl:301-322),   */
/* averagers (inc/average.jl:8-48, :52-97), re-init (mcmc_eap_chain.jl:352-361).               */
/* ------------------------------------------------------------------------------------------ */
struct orc_run {
  orc_case c;
  uint64_t seed;
  uint32_t chain_id;
  uint32_t init;
  int32_t algo;
  orc_chain* chain;
  orc_chain* trial;
  double logpi_prev;   /* Metropolis.logπ_prev (acceptance.jl:16) */
  double log_gauge;    /* AntiDipoleWeightFunction.log_gauge (average.jl:104-118) */
  double cF;           /* 0.2 + 0.8 exp(-(Fx²+Fz²)/kT) (average.jl:121-122) */
  double phi_step, theta_step;
  int64_t nacc, natt, nacc_total, steps_total;
  double acc[16];
  double normalizer;
  /* running quantities of the ΔU formulation (algo 1) */
  double p[3], su;
  /* clustering driver */
  double carry;          /* algo 1: logπ_prev − logπ(chain) = log α of the last accepted trial (acceptance.jl:30-33) */
  double acc_x[2];       /* Σ of the two extra averagers (mcmc_clustering_eap_chain.jl:243-244) */
  double spsi, scos2;    /* algo 1: running Σψ and Σcos²θ */
  double ncluster, cluster_sum, cluster_max;
  /* tape mode (see orc_run_new_tape) */
  const double* tape;
  int64_t tape_len, tape_pos;
  int32_t tape_underrun;
  int32_t reinit_stale;   /* 1 = the reference's literal behaviour: logπ_prev is NOT refreshed after a re-init swap */
};

/* Tape mode (tests only).  The reference draws from Julia's global RNG; to compare the oracle with the reference's
 * own code run on a scripted `rand` (tests/golden/ref_driver_*.jl), every random draw pops the next value of a
 * caller-supplied array of uniforms, in the ORDER the reference calls rand():
 *   EAPChain(pargs): n values for ϕ (ϕ = 2π·u), then n for θ (θ = π·u)                      eap_chain.jl:62
 *   trial: idx = ⌊u·n⌋, dϕ = −s + 2s·u, [--do-flips: Bool = u < ½], dθ                      mcmc_eap_chain.jl:277-280
 *          [cluster_flip!: 3-D gate, up-growth…, down-growth… / 2-D growth… then gate]      eap_chain.jl:269-333
 *          ϵ                                                                                mcmc_eap_chain.jl:287
 *   re-init: 2n values for the new chain, then ϵ unless --force-init                        mcmc_eap_chain.jl:352-357
 * The Markov chain logic is the SAME code as in Philox mode; only the source of the uniforms differs. */
static double tape_pop(orc_run* r) {
  if (r->tape_pos >= r->tape_len) {
    r->tape_underrun = 1;
    return 0.5;
  }
  return r->tape[r->tape_pos++];
}

typedef struct {
  int64_t idx;
  double uphi, uth, eps;
  int32_t flipbit;
} step_draws;

/* idx, dϕ, [Bool], dθ of trial `step` (mcmc_eap_chain.jl:277-280). */
static void draw_head(orc_run* r, int64_t step, step_draws* d) {
  const orc_case* c = &r->c;
  if (!r->tape) {
    orc_draw_step(r->seed, r->chain_id, r->init, step, c->n, &d->idx, &d->uphi, &d->flipbit, &d->uth, &d->eps);
    return;
  }
  d->idx = (int64_t)floor(tape_pop(r) * (double)c->n);
  d->uphi = tape_pop(r);
  d->flipbit = c->do_flips ? (tape_pop(r) < 0.5) : 0;   /* `pargs["do-flips"] && rand(Bool)` short-circuits */
  d->uth = c->planar ? 0.0 : tape_pop(r);               /* the 2-D driver draws no dθ */
  d->eps = 0.0;
}

/* ϵ of the acceptor call (mcmc_eap_chain.jl:287): drawn after everything else of the trial. */
static double draw_eps(orc_run* r, const step_draws* d) { return r->tape ? tape_pop(r) : d->eps; }

static void orc_run_steps_impl(orc_run* r, int64_t nsteps, int64_t stepout, double* traj, double* roll, int roll_cols,
                               double* state);

/* AntiDipoleWeightFunction (average.jl:109-124) or WeightlessFunction (=1.0, average.jl:102). */
static double weight_of(const orc_run* r, double su) {
  if (!r->c.umbrella) return 1.0;
  return su / r->c.kT * r->cF - r->log_gauge;
}

static void run_bind_chain(orc_run* r) {
  r->su = sum_us(r->chain);
  orc_chain_p(r->chain, r->p);
  r->spsi = sum_psi(r->chain);
  r->scos2 = sum_cos2(r->chain);
  r->carry = 0.0;
  r->logpi_prev = -r->chain->U / r->c.kT + r->chain->Omega + weight_of(r, r->su);
}

/* AntiDipoleWeightFunction(chain) (average.jl:109-118): the gauge is fixed when it is constructed. */
static void run_bind_gauge(orc_run* r) {
  const orc_case* c = &r->c;
  r->cF = 0.2 + 0.8 * exp(-(c->Fx * c->Fx + c->Fz * c->Fz) / c->kT);
  if (c->chain_type == ORC_CHAIN_DIELECTRIC)
    r->log_gauge = -(c->K1 + 2 * c->K2) * c->E0 * c->E0 * (double)c->n / (3 * c->kT) + r->chain->Omega;
  else /* eigvals(mu*I)[end] = mu */
    r->log_gauge = -c->mu * c->E0 * (double)c->n / (3 * c->kT) + r->chain->Omega;
}

/* EAPChain(pargs) on a tape: `rand(ϕ_dist, n), rand(θ_dist, n)` (eap_chain.jl:62) — all ϕ first, then all θ.  The
 * planar chain (2D/inc/eap_chain.jl:66) draws ϕ only. */
static orc_chain* chain_from_tape(orc_run* r) {
  const orc_case* c = &r->c;
  double* phi = (double*)malloc(sizeof(double) * (size_t)c->n);
  double* theta = (double*)malloc(sizeof(double) * (size_t)c->n);
  for (int64_t k = 0; k < c->n; ++k) phi[k] = 0.0 + (2.0 * M_PI - 0.0) * tape_pop(r);
  for (int64_t k = 0; k < c->n; ++k) theta[k] = c->planar ? 0.0 : 0.0 + (M_PI - 0.0) * tape_pop(r);
  orc_chain* ch = orc_chain_new(c, phi, theta);
  free(phi);
  free(theta);
  return ch;
}

orc_run* orc_run_new(const orc_case* c, uint64_t seed, uint32_t chain_id, int32_t algo) {
  orc_run* r = (orc_run*)calloc(1, sizeof(orc_run));
  r->c = *c;
  r->seed = seed;
  r->chain_id = chain_id;
  r->algo = algo;
  r->phi_step = c->phi_step;
  r->theta_step = c->theta_step;
  r->chain = orc_chain_new_random(c, seed, chain_id, 0);
  r->trial = orc_chain_copy(r->chain);
  run_bind_gauge(r);
  run_bind_chain(r);
  return r;
}

/* A fresh `mcmc(nsteps, pargs, chain)` call of the clustering driver on the current chain
 * (mcmc_clustering_eap_chain.jl:171-265): chain.kT = kT (:172), chain.U = U(chain) (:176), new weight
 * function (:177) and acceptor (:178-181), new averagers (:196-251), counters and step sizes (:173,:262-264).
 * burnargs is a *copy* of pargs (:368,:378), so the adapted step sizes saved at :348-349 never reach the
 * next stage: every stage starts from --phi-step/--theta-step. */
void orc_run_begin_stage(orc_run* r, double kT) {
  r->c.kT = kT;
  r->chain->c.kT = kT;
  r->trial->c.kT = kT;
  r->chain->U = U_total(r->chain);
  r->init += 1; /* fresh random numbers for the new stage */
  if (r->c.planar) { /* 2D/mcmc_clustering_eap_chain.jl:151 `chain = EAPChain(pargs)`: every stage builds a NEW chain */
    orc_chain* nc = r->tape ? chain_from_tape(r) : orc_chain_new_random(&r->c, r->seed, r->chain_id, r->init);
    chain_assign(r->chain, nc);
    orc_chain_free(nc);
  }
  r->phi_step = r->c.phi_step;
  r->theta_step = r->c.theta_step;
  r->nacc = r->natt = r->nacc_total = r->steps_total = 0;
  memset(r->acc, 0, sizeof(r->acc));
  memset(r->acc_x, 0, sizeof(r->acc_x));
  r->normalizer = 0.0;
  r->ncluster = r->cluster_sum = r->cluster_max = 0.0;
  run_bind_gauge(r);
  run_bind_chain(r);
}

void orc_run_init_x0(orc_run* r, const double* x0, int64_t x0_len, const double dx0[2]) {
  orc_chain* ch;
  if (r->tape) {
    /* EAPChain(pargs) with --x0 draws only the perturbations: rand(Uniform(0, dx0[1]), n), then rand(Uniform(0, dx0[2]), n)
     * (eap_chain.jl:70-75) — the tape restarts, the random chain of the constructor was never drawn by the reference */
    if (x0_len != 2 && x0_len != 2 * r->c.n) return;
    r->tape_pos = 0;
    const int64_t n = r->c.n;
    double* phi = (double*)malloc(sizeof(double) * (size_t)n);
    double* theta = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t k = 0; k < n; ++k) phi[k] = (x0_len == 2 ? x0[0] : x0[2 * k]) + (0.0 + (dx0[0] - 0.0) * tape_pop(r));
    for (int64_t k = 0; k < n; ++k) theta[k] = (x0_len == 2 ? x0[1] : x0[2 * k + 1]) + (0.0 + (dx0[1] - 0.0) * tape_pop(r));
    ch = orc_chain_new(&r->c, phi, theta);
    free(phi);
    free(theta);
  } else {
    ch = orc_chain_new_x0(&r->c, r->seed, r->chain_id, 0, x0, x0_len, dx0);
  }
  if (!ch) return;
  chain_assign(r->chain, ch);
  orc_chain_free(ch);
  run_bind_gauge(r);
  run_bind_chain(r);
}

void orc_run_set_state(orc_run* r, const double* phi, const double* theta) {
  orc_chain* ch = orc_chain_new(&r->c, phi, theta);
  chain_assign(r->chain, ch);
  orc_chain_free(ch);
  run_bind_chain(r);
}

/* record! (average.jl:40-48 standard; :63-73 umbrella): 8 averagers = 16 sums + normaliser,
 * built at mcmc_eap_chain.jl:243-255, recorded every trial incl. rejected (:327-328). */
static void record(orc_run* r) {
  const orc_chain* ch = r->chain;
  double rr[3], p[3];
  if (r->algo == 0) {
    end_to_end(ch, rr);
    orc_chain_p(ch, p);
  } else {
    memcpy(rr, ch->r, sizeof(rr));
    memcpy(p, r->p, sizeof(p));
  }
  double v[16] = {rr[0], rr[1], rr[2], rr[0] * rr[0], rr[1] * rr[1], rr[2] * rr[2],
                  rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2],
                  p[0], p[1], p[2], p[0] * p[0], p[1] * p[1], p[2] * p[2],
                  p[0] * p[0] + p[1] * p[1] + p[2] * p[2],
                  ch->U, ch->U * ch->U};
  /* mcmc_clustering_eap_chain.jl:243-244: Σcos²θ and Σψ/(n−1) */
  double x[2] = {(r->algo == 0) ? sum_cos2(ch) : r->scos2,
                 ((r->algo == 0) ? sum_psi(ch) : r->spsi) / (double)(ch->n - 1)};
  if (r->c.umbrella) {
    double su = (r->algo == 0) ? sum_us(ch) : r->su;
    double expw = exp(weight_of(r, su));
    for (int k = 0; k < 16; ++k) r->acc[k] += v[k] / expw;
    for (int k = 0; k < 2; ++k) r->acc_x[k] += x[k] / expw;
    r->normalizer += 1.0 / expw;
  } else {
    for (int k = 0; k < 16; ++k) r->acc[k] += v[k];
    for (int k = 0; k < 2; ++k) r->acc_x[k] += x[k];
    r->normalizer += 1;
  }
}

static void emit_rows(const orc_run* r, int64_t step, double* traj_row, double* roll_row, int roll_cols,
                      double* state_row) {
  const orc_chain* ch = r->chain;
  if (traj_row) { /* mcmc_eap_chain.jl:330-333 */
    double p[3];
    if (r->algo == 0) orc_chain_p(ch, p); else memcpy(p, r->p, sizeof(p));
    traj_row[0] = (double)step;
    traj_row[1] = ch->r[0]; traj_row[2] = ch->r[1]; traj_row[3] = ch->r[2];
    traj_row[4] = p[0]; traj_row[5] = p[1]; traj_row[6] = p[2];
    traj_row[7] = ch->U;
  }
  if (roll_row) { /* mcmc_eap_chain.jl:334-346 */
    roll_row[0] = (double)step;
    for (int k = 0; k < 16; ++k) roll_row[1 + k] = r->acc[k] / r->normalizer;
    if (roll_cols == 19) /* mcmc_clustering_eap_chain.jl:344-345 */
      for (int k = 0; k < 2; ++k) roll_row[17 + k] = r->acc_x[k] / r->normalizer;
  }
  if (state_row) /* :317: phi1,theta1,phi2,theta2,… (the μ columns :318 are a function of these) */
    for (int64_t i = 0; i < ch->n; ++i) {
      state_row[2 * i] = ch->phi[i];
      state_row[2 * i + 1] = ch->theta[i];
    }
}

/* cluster_flip! (eap_chain.jl:269-333) on the chain `t` that already carries the single-monomer move:
 * decides the cluster [lo,hi] and the link probabilities at its ends.  Returns 0 if the gate draw says
 * "no cluster" (rand() <= ϵflip → α = 1, :273). */
static int cluster_grow(orc_run* r, const orc_chain* t, int64_t step, int64_t idx, int64_t* lo, int64_t* hi,
                        double* upper_p, double* lower_p) {
  const orc_case* c = &r->c;
  if (!c->clustering) return 0; /* mcmc_eap_chain.jl has no cluster_flip! */
  /* 3-D: the gate comes first and a hit means "no cluster" (eap_chain.jl:273).  2-D: the cluster is grown
   * first and a hit means "flip it" (2D/inc/eap_chain.jl:233).  The Philox growth uniforms are counter-based, so
   * growing after the gate changes nothing there; on a tape the calls happen in the reference's order. */
  if (!c->planar) {
    const double g = r->tape ? tape_pop(r) : orc_draw_cluster_gate(r->seed, r->chain_id, r->init, step);
    if (g <= c->cluster_prob) return 0;
  } else if (!r->tape) {
    if (!(orc_draw_cluster_gate(r->seed, r->chain_id, r->init, step) <= c->cluster_prob)) return 0;
  }
  int64_t u = idx, k = 0;
  double up;
  for (;;) { /* :276-289 */
    if (u >= t->n - 1) { up = 0.0; break; }
    up = orc_chain_link_prob(t, u);
    const double x = r->tape ? tape_pop(r) : orc_draw_cluster(r->seed, r->chain_id, r->init, step, 0, k++);
    if (x <= up) u += 1; else break;
  }
  int64_t l = idx;
  double lp;
  k = 0;
  for (;;) { /* :292-305 */
    if (l <= 0) { lp = 0.0; break; }
    lp = orc_chain_link_prob(t, l - 1);
    const double x = r->tape ? tape_pop(r) : orc_draw_cluster(r->seed, r->chain_id, r->init, step, 1, k++);
    if (x <= lp) l -= 1; else break;
  }
  if (c->planar && r->tape && !(tape_pop(r) <= c->cluster_prob)) return 0;  /* 2D/inc/eap_chain.jl:233 */
  *lo = l; *hi = u; *upper_p = up; *lower_p = lp;
  return 1;
}

void orc_run_steps(orc_run* r, int64_t nsteps, int64_t stepout, double* traj, double* roll) {
  orc_run_steps_impl(r, nsteps, stepout, traj, roll, 17, NULL);
}

void orc_run_steps_ex(orc_run* r, int64_t nsteps, int64_t stepout, double* traj, double* roll19, double* state) {
  orc_run_steps_impl(r, nsteps, stepout, traj, roll19, 19, state);
}

/* One trial of mcmc_clustering_eap_chain.jl:267-279: move!, cluster_flip!, acceptor with α. */
static int cluster_trial(orc_run* r, int64_t step) {
  const orc_case* c = &r->c;
  step_draws sd;
  draw_head(r, step, &sd);
  const int64_t idx = sd.idx;
  const double uphi = sd.uphi, uth = sd.uth;
  const int32_t flipbit = sd.flipbit;
  const double dphi = -r->phi_step + (2 * r->phi_step) * uphi;   /* :268-270 */
  /* (--do-flips exists only in mcmc_eap_chain.jl:279, which reaches this path when it carries bending energy) */
  const double dth = c->planar ? 0.0 /* no dθ draw in 2D/mcmc_clustering_eap_chain.jl:238-241 */
                               : ((c->do_flips && flipbit) ? M_PI - 2 * r->chain->theta[idx] : 0.0) +
                                     (-r->theta_step + (2 * r->theta_step) * uth);
  int64_t lo = idx, hi = idx;
  double up = 0.0, lp = 0.0, alpha = 1.0;
  int reflect = 0, accepted = 0;
  if (r->algo == 0) {
    chain_assign(r->trial, r->chain);                 /* :271 */
    orc_chain_move(r->trial, idx, dphi, dth);         /* :272 */
    reflect = cluster_grow(r, r->trial, step, idx, &lo, &hi, &up, &lp);
    const double eps = draw_eps(r, &sd);              /* :274 `acceptor(trial_chain, rand(); α = α)` */
    if (reflect) {
      for (int64_t i = lo; i <= hi; ++i) {            /* eap_chain.jl:311-315 */
        chain_refl_caches(r->trial, i);
        r->trial->U = U_total(r->trial);
      }
      const double nup = hi < c->n - 1 ? orc_chain_link_prob(r->trial, hi) : 0.0;      /* :317-321 */
      const double nlp = lo > 0 ? orc_chain_link_prob(r->trial, lo - 1) : 0.0;         /* :322-326 */
      alpha = ((1 - nup) * (1 - nlp)) / ((1 - up) * (1 - lp));                         /* :327-330 */
    }
    /* acceptance.jl:29-37 */
    const double logpi_chain = -r->trial->U / c->kT + r->trial->Omega + weight_of(r, c->umbrella ? sum_us(r->trial) : 0.0);
    const double logpi = logpi_chain + log(alpha);
    if ((logpi >= r->logpi_prev) || (eps < exp(logpi - r->logpi_prev))) {
      r->logpi_prev = c->alpha_carry ? logpi : logpi_chain;
      orc_chain* t = r->chain; r->chain = r->trial; r->trial = t;
      accepted = 1;
    }
  } else {
    /* link probabilities are read from the chain carrying the single-monomer move: build only n̂'_idx */
    orc_chain* t = r->trial; /* scratch view: copy n̂ of idx's neighbourhood lazily = copy all n̂ (cheap, O(n)) */
    memcpy(t->nhat, r->chain->nhat, sizeof(double) * 3 * (size_t)c->n);
    {
      const double phi1 = r->chain->phi[idx] + dphi;
      const double th1 = fmin(M_PI, fmax(0.0, r->chain->theta[idx] + dth));
      if (c->planar) nhat_planar(cos(phi1), sin(phi1), &t->nhat[3 * idx]);
      else nhat_of(cos(phi1), sin(phi1), cos(th1), sin(th1), &t->nhat[3 * idx]);
    }
    reflect = cluster_grow(r, t, step, idx, &lo, &hi, &up, &lp);
    const double eps = draw_eps(r, &sd);
    double d[12];
    orc_chain_delta_segment(r->chain, idx, dphi, dth, reflect, lo, hi, d);
    if (reflect) {
      /* n̂ after the reflections at the two ends of the cluster: refl_n! maps θ → π−θ */
      for (int64_t i = lo; i <= hi; i += (hi > lo ? hi - lo : 1)) {
        double phi1 = r->chain->phi[i], th1 = r->chain->theta[i];
        if (i == idx) { phi1 += dphi; th1 = fmin(M_PI, fmax(0.0, th1 + dth)); }
        if (c->planar) {
          phi1 += M_PI;
          nhat_planar(cos(phi1), sin(phi1), &t->nhat[3 * i]);
          continue;
        }
        th1 = fmin(M_PI, fmax(0.0, th1 + (M_PI - 2 * th1)));
        nhat_of(cos(phi1), sin(phi1), cos(th1), sin(th1), &t->nhat[3 * i]);
      }
      const double nup = hi < c->n - 1 ? orc_chain_link_prob(t, hi) : 0.0;
      const double nlp = lo > 0 ? orc_chain_link_prob(t, lo - 1) : 0.0;
      alpha = ((1 - nup) * (1 - nlp)) / ((1 - up) * (1 - lp));
    }
    const double dsu = d[3] + d[6];
    const double dw = c->umbrella ? dsu / c->kT * r->cF : 0.0;
    const double la = log(alpha);
    const double dlogpi = -d[0] / c->kT + d[1] + dw + la - r->carry;
    if ((dlogpi >= 0.0) || (eps < exp(dlogpi))) {
      orc_chain* ch = r->chain;
      const double Unew = ch->U + d[0], Onew = ch->Omega + d[1];
      chain_move_caches(ch, idx, dphi, dth);
      if (reflect)
        for (int64_t i = lo; i <= hi; ++i) chain_refl_caches(ch, i);
      ch->U = Unew;
      ch->Omega = Onew;
      for (int k = 0; k < 3; ++k) r->p[k] += d[9 + k];
      r->su += dsu;
      r->spsi += d[7];
      r->scos2 += d[8];
      r->carry = c->alpha_carry ? la : 0.0;
      accepted = 1;
    }
  }
  if (reflect) {
    const double sz = (double)(hi - lo + 1);
    r->ncluster += 1; r->cluster_sum += sz;
    if (sz > r->cluster_max) r->cluster_max = sz;
  }
  return accepted;
}

static void orc_run_steps_impl(orc_run* r, int64_t nsteps, int64_t stepout, double* traj, double* roll, int roll_cols,
                               double* state) {
  const orc_case* c = &r->c;
  int64_t row = 0;
  for (int64_t step = 1; step <= nsteps; ++step) {
    if (c->clustering || c->kappa != 0.0 || c->energy_type == ORC_ENERGY_CUTOFF) {
      const int acc = cluster_trial(r, step);
      if (acc) { r->nacc += 1; r->nacc_total += 1; }
      goto counted;
    }
    {
    step_draws sd;
    draw_head(r, step, &sd);
    const int64_t idx = sd.idx;
    const double uphi = sd.uphi, uth = sd.uth, eps = draw_eps(r, &sd);
    const int32_t flipbit = sd.flipbit;
    /* rand(Uniform(-s, s)) = -s + 2s*u  (mcmc_eap_chain.jl:278-280) */
    double dphi = -r->phi_step + (2 * r->phi_step) * uphi;
    double dth = ((c->do_flips && flipbit) ? M_PI - 2 * r->chain->theta[idx] : 0.0) +
                 (-r->theta_step + (2 * r->theta_step) * uth);
    int accepted = 0;
    if (r->algo == 0) {
      /* trial = deep copy; move!; full U; stateful acceptor (mcmc_eap_chain.jl:281-292) */
      chain_assign(r->trial, r->chain);
      orc_chain_move(r->trial, idx, dphi, dth);
      double logpi = -r->trial->U / c->kT + r->trial->Omega + weight_of(r, c->umbrella ? sum_us(r->trial) : 0.0);
      if ((logpi >= r->logpi_prev) || (eps < exp(logpi - r->logpi_prev))) {
        r->logpi_prev = logpi;
        orc_chain* t = r->chain; r->chain = r->trial; r->trial = t;
        accepted = 1;
      }
    } else {
      double d[6];
      orc_chain_delta_u(r->chain, idx, dphi, dth, d);
      double dw = c->umbrella ? d[3] / c->kT * r->cF : 0.0;
      double dlogpi = -d[0] / c->kT + d[1] + dw;
      if ((dlogpi >= 0.0) || (eps < exp(dlogpi))) {
        orc_chain* ch = r->chain;
        double old_mu[3] = {ch->mus[3 * idx], ch->mus[3 * idx + 1], ch->mus[3 * idx + 2]};
        double Unew = ch->U + d[0], Onew = ch->Omega + d[1];
        chain_move_caches(ch, idx, dphi, dth);
        ch->U = Unew;
        ch->Omega = Onew;
        for (int k = 0; k < 3; ++k) r->p[k] += ch->mus[3 * idx + k] - old_mu[k];
        r->su += d[3];
        accepted = 1;
      }
    }
    if (accepted) { r->nacc += 1; r->nacc_total += 1; }
    }
  counted:
    r->natt += 1;
    r->steps_total += 1;
    /* step-size adaptation (mcmc_eap_chain.jl:301-322) */
    if (c->adj_scale != 1.0 && c->steps_per_adjust > 0 && step % c->steps_per_adjust == 0) {
      double ratio = (double)r->nacc / (double)r->natt;
      if (ratio > c->adj_ub && r->phi_step != M_PI && r->theta_step != M_PI / 2) {
        r->nacc = 0; r->natt = 0;
        r->phi_step = fmin(M_PI, r->phi_step * c->adj_scale);
        r->theta_step = fmin(M_PI / 2, r->theta_step * c->adj_scale);
      } else if (ratio < c->adj_lb) {
        r->nacc = 0; r->natt = 0;
        r->phi_step /= c->adj_scale;
        r->theta_step /= c->adj_scale;
      }
    }
    record(r);
    if (stepout > 0 && step % stepout == 0) {
      emit_rows(r, step, traj ? traj + 8 * row : NULL, roll ? roll + roll_cols * row : NULL, roll_cols,
                state ? state + 2 * c->n * row : NULL);
      ++row;
    }
  }
}

/* Re-initialise between inits (mcmc_eap_chain.jl:352-361; metropolis_acc acceptance.jl:1-3):
 * accept iff force || eps <= exp(-dU/kT) * Π sinθ_new / Π sinθ_old.  The reference leaves the
 * acceptor's logπ_prev stale after the swap (Appendix B); fixed here by rebinding. */
int32_t orc_run_reinit(orc_run* r, int32_t force_init) {
  r->init += 1;
  orc_chain* nc;
  double eps = 0.0;
  if (r->tape) {
    nc = chain_from_tape(r);
    if (!force_init) eps = tape_pop(r);   /* `pargs["force-init"] || metropolis_acc(..., rand())` short-circuits */
  } else {
    nc = orc_chain_new_random(&r->c, r->seed, r->chain_id, r->init);
    eps = orc_draw_reinit_eps(r->seed, r->chain_id, r->init);
  }
  int32_t take = force_init || (eps <= exp(-(nc->U - r->chain->U) / r->c.kT + (nc->Omega - r->chain->Omega)));
  if (take) {
    const double stale = r->logpi_prev;
    chain_assign(r->chain, nc);
    run_bind_chain(r);
    if (r->reinit_stale) r->logpi_prev = stale;   /* mcmc_eap_chain.jl:360: `chain = new_chain`, acceptor untouched */
  }
  orc_chain_free(nc);
  return take;
}

/* Tape mode constructor: the initial chain comes from the tape too. */
orc_run* orc_run_new_tape(const orc_case* c, int32_t algo, const double* tape, int64_t tape_len, int32_t reinit_stale) {
  orc_run* r = (orc_run*)calloc(1, sizeof(orc_run));
  r->c = *c;
  r->algo = algo;
  r->phi_step = c->phi_step;
  r->theta_step = c->theta_step;
  r->tape = tape;
  r->tape_len = tape_len;
  r->reinit_stale = reinit_stale;
  r->chain = chain_from_tape(r);
  r->trial = orc_chain_copy(r->chain);
  run_bind_gauge(r);
  run_bind_chain(r);
  return r;
}

int64_t orc_run_tape_pos(const orc_run* r) { return r->tape_underrun ? -1 : r->tape_pos; }

void orc_run_extra_averages(const orc_run* r, double ex[2]) {
  for (int k = 0; k < 2; ++k) ex[k] = r->acc_x[k] / r->normalizer;
}

void orc_run_cluster_stats(const orc_run* r, double out[3]) {
  out[0] = r->ncluster; out[1] = r->cluster_sum; out[2] = r->cluster_max;
}

void orc_run_averages(const orc_run* r, double avg[16], double* acc_rate, double* normalizer) {
  for (int k = 0; k < 16; ++k) avg[k] = r->acc[k] / r->normalizer;
  if (acc_rate) *acc_rate = r->steps_total ? (double)r->nacc_total / (double)r->steps_total : 0.0;
  if (normalizer) *normalizer = r->normalizer;
}

void orc_run_diag(const orc_run* r, double out[8]) {
  out[0] = r->phi_step; out[1] = r->theta_step;
  out[2] = (double)r->nacc; out[3] = (double)r->natt;
  out[4] = (double)r->nacc_total; out[5] = (double)r->steps_total;
  out[6] = r->chain->U; out[7] = r->chain->Omega;
}

const orc_chain* orc_run_chain(const orc_run* r) { return r->chain; }

void orc_run_free(orc_run* r) {
  if (!r) return;
  orc_chain_free(r->chain);
  orc_chain_free(r->trial);
  free(r);
}

/* ------------------------------------------------------------------------------------------ */
/* Throughput probe: the reference runs one chain per OS process (run/ launchers, `julia -t 1`); the  */
/* stand-in is one chain per pthread.                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const orc_case* c;
  uint64_t seed;
  int32_t algo;
  int32_t first, last;
  int64_t nsteps;
  double sink;
} bench_job;

static void* bench_worker(void* arg) {
  bench_job* j = (bench_job*)arg;
  for (int32_t k = j->first; k < j->last; ++k) {
    orc_run* r = orc_run_new(j->c, j->seed, (uint32_t)k, j->algo);
    orc_run_steps(r, j->nsteps, 0, NULL, NULL);
    j->sink += r->acc[14];
    orc_run_free(r);
  }
  return NULL;
}

double orc_bench(const orc_case* c, uint64_t seed, int32_t algo, int32_t nchains, int64_t nsteps, int32_t nthreads) {
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  bench_job* jobs = (bench_job*)calloc((size_t)nthreads, sizeof(bench_job));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int32_t t = 0; t < nthreads; ++t) {
    jobs[t].c = c; jobs[t].seed = seed; jobs[t].algo = algo; jobs[t].nsteps = nsteps;
    jobs[t].first = (int32_t)((int64_t)nchains * t / nthreads);
    jobs[t].last = (int32_t)((int64_t)nchains * (t + 1) / nthreads);
    pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
  }
  for (int32_t t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th);
  free(jobs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
