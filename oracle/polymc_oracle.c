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
enum { SUB_STEP_A = 0, SUB_STEP_B = 1, SUB_INIT = 2, SUB_REINIT = 3 };

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

double orc_draw_reinit_eps(uint64_t seed, uint32_t chain_id, uint32_t init) {
  uint32_t w[4];
  philox_at(seed, chain_id, init, SUB_REINIT, 0, w);
  return u53(w[0], w[1]);
}

/* ------------------------------------------------------------------------------------------ */
/* EAPChain (inc/eap_chain.jl:12-36).  kappa == 0 in this driver (no --bend-mod; :91), so the   */
/* psi / ubend caches (:45-47,:54-58) carry no energy and are not stored.                       */
/* ------------------------------------------------------------------------------------------ */
struct orc_chain {
  orc_case c;
  int64_t n;
  double *phi, *cphi, *sphi, *theta, *ctheta, *stheta;
  double *nhat, *mus, *us, *xs; /* 3n, 3n, n, 3n */
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

/* u(E0, mu) = -1/2 E0 mu_z  (eap_chain.jl:53); same 1/2 for polar chains. */
static inline double u_self(double E0, const double mu[3]) { return -1.0 / 2 * E0 * mu[2]; }

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
  return 0.0;
}

/* Energy functors (inc/energy.jl:7-9, :13-16, :20-23). */
static double U_total(const orc_chain* ch) {
  double r[3];
  end_to_end(ch, r);
  double rf = r[0] * ch->c.Fx + r[1] * 0.0 + r[2] * ch->c.Fz;
  if (ch->c.energy_type == ORC_ENERGY_NONINTERACTING) return sum_us(ch) - rf;
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
    prod *= ch->stheta[i];
    slog += log(ch->stheta[i]);
    nhat_of(ch->cphi[i], ch->sphi[i], ch->ctheta[i], ch->stheta[i], &ch->nhat[3 * i]);
    mu_of(c, ch->cphi[i], ch->sphi[i], ch->ctheta[i], ch->stheta[i], &ch->mus[3 * i]);
    ch->us[i] = u_self(c->E0, &ch->mus[3 * i]);
  }
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
  memcpy(d->r, s->r, sizeof(d->r));
  d->Omega = s->Omega;
  d->U = s->U;
}

void orc_chain_free(orc_chain* ch) {
  if (!ch) return;
  free(ch->phi); free(ch->cphi); free(ch->sphi);
  free(ch->theta); free(ch->ctheta); free(ch->stheta);
  free(ch->nhat); free(ch->mus); free(ch->us); free(ch->xs);
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
  if (ch->c.energy_type == ORC_ENERGY_INTERACTING) {
    for (int64_t i = 0; i < ch->n; ++i)
      for (int64_t j = i + 1; j < ch->n; ++j)
        s += fabs(pair_term(&ch->xs[3 * i], &ch->xs[3 * j], &ch->mus[3 * i], &ch->mus[3 * j]));
  } else if (ch->c.energy_type == ORC_ENERGY_ISING) {
    for (int64_t i = 0; i + 1 < ch->n; ++i)
      s += fabs(pair_term(&ch->xs[3 * i], &ch->xs[3 * (i + 1)], &ch->mus[3 * i], &ch->mus[3 * (i + 1)]));
  }
  return s;
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
  ch->theta[idx] = fmin(M_PI, fmax(0.0, ch->theta[idx] + dtheta));
  double sth = sin(ch->theta[idx]);
  ch->Omega += log(sth / ch->stheta[idx]);
  ch->ctheta[idx] = cos(ch->theta[idx]);
  ch->stheta[idx] = sth;
  nhat_of(ch->cphi[idx], ch->sphi[idx], ch->ctheta[idx], ch->stheta[idx], &ch->nhat[3 * idx]);
  mu_of(&ch->c, ch->cphi[idx], ch->sphi[idx], ch->ctheta[idx], ch->stheta[idx], &ch->mus[3 * idx]);
  if (idx > 0) ch->us[idx - 1] = u_self(ch->c.E0, &ch->mus[3 * (idx - 1)]);
  ch->us[idx] = u_self(ch->c.E0, &ch->mus[3 * idx]);
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
  double du = u_self(c->E0, mu1) - ch->us[idx];
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
};

/* AntiDipoleWeightFunction (average.jl:109-124) or WeightlessFunction (=1.0, average.jl:102). */
static double weight_of(const orc_run* r, double su) {
  if (!r->c.umbrella) return 1.0;
  return su / r->c.kT * r->cF - r->log_gauge;
}

static void run_bind_chain(orc_run* r) {
  r->su = sum_us(r->chain);
  orc_chain_p(r->chain, r->p);
  r->logpi_prev = -r->chain->U / r->c.kT + r->chain->Omega + weight_of(r, r->su);
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
  r->cF = 0.2 + 0.8 * exp(-(c->Fx * c->Fx + c->Fz * c->Fz) / c->kT);
  if (c->chain_type == ORC_CHAIN_DIELECTRIC)
    r->log_gauge = -(c->K1 + 2 * c->K2) * c->E0 * c->E0 * (double)c->n / (3 * c->kT) + r->chain->Omega;
  else /* eigvals(mu*I)[end] = mu */
    r->log_gauge = -c->mu * c->E0 * (double)c->n / (3 * c->kT) + r->chain->Omega;
  run_bind_chain(r);
  return r;
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
  if (r->c.umbrella) {
    double su = (r->algo == 0) ? sum_us(ch) : r->su;
    double expw = exp(weight_of(r, su));
    for (int k = 0; k < 16; ++k) r->acc[k] += v[k] / expw;
    r->normalizer += 1.0 / expw;
  } else {
    for (int k = 0; k < 16; ++k) r->acc[k] += v[k];
    r->normalizer += 1;
  }
}

static void emit_rows(const orc_run* r, int64_t step, double* traj_row, double* roll_row) {
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
  }
}

void orc_run_steps(orc_run* r, int64_t nsteps, int64_t stepout, double* traj, double* roll) {
  const orc_case* c = &r->c;
  int64_t row = 0;
  for (int64_t step = 1; step <= nsteps; ++step) {
    int64_t idx;
    double uphi, uth, eps;
    int32_t flipbit;
    orc_draw_step(r->seed, r->chain_id, r->init, step, c->n, &idx, &uphi, &flipbit, &uth, &eps);
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
      emit_rows(r, step, traj ? traj + 8 * row : NULL, roll ? roll + 17 * row : NULL);
      ++row;
    }
  }
}

/* Re-initialise between inits (mcmc_eap_chain.jl:352-361; metropolis_acc acceptance.jl:1-3):
 * accept iff force || eps <= exp(-dU/kT) * Π sinθ_new / Π sinθ_old.  The reference leaves the
 * acceptor's logπ_prev stale after the swap (Appendix B); fixed here by rebinding. */
int32_t orc_run_reinit(orc_run* r, int32_t force_init) {
  r->init += 1;
  orc_chain* nc = orc_chain_new_random(&r->c, r->seed, r->chain_id, r->init);
  double eps = orc_draw_reinit_eps(r->seed, r->chain_id, r->init);
  int32_t take = force_init || (eps <= exp(-(nc->U - r->chain->U) / r->c.kT + (nc->Omega - r->chain->Omega)));
  if (take) {
    chain_assign(r->chain, nc);
    run_bind_chain(r);
  }
  orc_chain_free(nc);
  return take;
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
