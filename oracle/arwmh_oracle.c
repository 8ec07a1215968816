/* CPU oracle in C (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the reference's ARWMH hot path, chain-parallel with
 * OpenMP.  It exists to (a) cross-check oracle/arwmh_numpy.py, (b) give parity
 * tests an oracle that finishes long trajectories in seconds, and (c) serve as
 * the `cpu_baseline` / `--impl reference` timing arm of bench.py (kind "port":
 * the real reference needs JAX + NumPyro, absent from this image).
 * PARITY STATUS: parity unpinned (see oracle/arwmh_numpy.py header).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py may load this library.
 *
 * This file is compiled twice (-DREAL_IS_DOUBLE=0/1) into one shared object
 * by oracle/Makefile; symbols carry an _f32 / _f64 suffix.
 *
 * Citations: python/kernels/arwmh.py (step :140-207), NumPyro cholesky_update
 * (3rd-party, restated), model files under python/scripts/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if REAL_IS_DOUBLE
typedef double real;
#define SUF(name) name##_f64
#define R_EXP exp
#define R_LOG log
#define R_LOG1P log1p
#define R_SQRT sqrt
#define R_POW pow
#else
typedef float real;
#define SUF(name) name##_f32
#define R_EXP expf
#define R_LOG logf
#define R_LOG1P log1pf
#define R_SQRT sqrtf
#define R_POW powf
#endif

enum { M_STD_NORMAL = 0, M_EIGHT_SCHOOLS = 1, M_KIDIQ = 2, M_DIAMONDS = 3, M_GAUSSIAN = 4 };

#define LOG_2PI_HALF 0.91893853320467274178

typedef struct {
  int model;
  int d;
  int64_t n;          /* data rows (kidiq, diamonds) */
  const real *a0, *a1, *a2; /* model arrays converted to `real` */
} SUF(model_t);

/* ---- potentials ------------------------------------------------------- */
/* python/scripts/run_eight_schools_lr_decay.py:26-35 */
static real SUF(pot_eight_schools)(const SUF(model_t) * m, const real *q) {
  const real *y = m->a0, *sg = m->a1;
  real mu = q[0], t = q[1];
  real tau = R_EXP(t);
  real lp_mu = -(real)0.5 * (mu / (real)5) * (mu / (real)5) - (real)(1.6094379124341003 + LOG_2PI_HALF);
  real tq = tau / (real)5;
  real lp_tau = (real)(0.6931471805599453 - 1.1447298858494002 - 1.6094379124341003) - R_LOG1P(tq * tq) + t;
  real lp_eta = 0, lp_obs = 0;
  for (int j = 0; j < 8; ++j) {
    real e = q[2 + j];
    lp_eta += -(real)0.5 * e * e - (real)LOG_2PI_HALF;
    real theta = mu + tau * e;
    real r = (y[j] - theta) / sg[j];
    lp_obs += -(real)0.5 * r * r - R_LOG(sg[j]) - (real)LOG_2PI_HALF;
  }
  return -(lp_mu + lp_tau + lp_eta + lp_obs);
}

static real SUF(student_t_logpdf)(real x, double df, double loc, double scale) {
  real yv = (x - (real)loc) / (real)scale;
  double zc = log(scale) + 0.5 * log(df) + 0.5 * log(M_PI) + lgamma(0.5 * df) - lgamma(0.5 * (df + 1.0));
  return -(real)(0.5 * (df + 1.0)) * R_LOG1P(yv * yv / (real)df) - (real)zc;
}

/* python/scripts/run_diamonds_lr_decay.py:24-40; a0 = Xc [n, Kc] row-major (centred), a1 = Y[n] */
static real SUF(pot_diamonds)(const SUF(model_t) * m, const real *q) {
  const int Kc = m->d - 2;
  const real *Xc = m->a0, *Y = m->a1;
  real icpt = q[0], s = q[1 + Kc];
  const real *b = q + 1;
  real sig = R_EXP(s);
  real lp_b = 0;
  for (int k = 0; k < Kc; ++k) lp_b += -(real)0.5 * b[k] * b[k] - (real)LOG_2PI_HALF;
  real lp_i = SUF(student_t_logpdf)(icpt, 3.0, 8.0, 10.0);
  real lp_s = (real)0.6931471805599453 + SUF(student_t_logpdf)(sig, 3.0, 0.0, 10.0) + s;
  real ss = 0;
  for (int64_t n = 0; n < m->n; ++n) {
    real mu = 0;
    const real *xr = Xc + n * Kc;
    for (int k = 0; k < Kc; ++k) mu += xr[k] * b[k];
    real r = (Y[n] - (icpt + mu)) / sig;
    ss += -(real)0.5 * r * r;
  }
  real lp_y = ss - (real)m->n * (s + (real)LOG_2PI_HALF);
  return -(lp_b + lp_i + lp_s + lp_y);
}

/* python/scripts/run_kidiq_kidscore_lr_decay.py:29-41; a0 = kid_score, a1 = mom_hs, a2 = mom_iq */
static real SUF(pot_kidiq)(const SUF(model_t) * m, const real *q) {
  real b0 = q[0], b1 = q[1], b2 = q[2], s = q[3];
  real sig = R_EXP(s);
  real sq = sig / (real)2.5;
  real lp_s = (real)(0.6931471805599453 - 1.1447298858494002 - 0.9162907318741551) - R_LOG1P(sq * sq) + s;
  real ss = 0;
  for (int64_t n = 0; n < m->n; ++n) {
    real mu = b0 + b1 * m->a1[n] + b2 * m->a2[n];
    real r = (m->a0[n] - mu) / sig;
    ss += -(real)0.5 * r * r;
  }
  real lp_y = ss - (real)m->n * (s + (real)LOG_2PI_HALF);
  return -(lp_s + lp_y);
}

static real SUF(pot_std_normal)(const SUF(model_t) * m, const real *q) {
  real ss = 0;
  for (int k = 0; k < m->d; ++k) ss += q[k] * q[k];
  return (real)0.5 * ss;
}

/* BASELINE.json config 5: a0 = P [d,d] row-major lower Cholesky of the precision; U = 0.5 |P^T q|^2 */
static real SUF(pot_gaussian)(const SUF(model_t) * m, const real *q) {
  const int d = m->d;
  real ss = 0;
  for (int j = 0; j < d; ++j) {
    real v = 0;
    for (int i = j; i < d; ++i) v += q[i] * m->a0[(int64_t)i * d + j];
    ss += v * v;
  }
  return (real)0.5 * ss;
}

static real SUF(potential)(const SUF(model_t) * m, const real *q) {
  switch (m->model) {
    case M_EIGHT_SCHOOLS: return SUF(pot_eight_schools)(m, q);
    case M_DIAMONDS: return SUF(pot_diamonds)(m, q);
    case M_KIDIQ: return SUF(pot_kidiq)(m, q);
    case M_GAUSSIAN: return SUF(pot_gaussian)(m, q);
    default: return SUF(pot_std_normal)(m, q);
  }
}

/* ---- cholesky_update (3rd-party NumPyro, restated; see arwmh_numpy.py) ---- */
/* L [d,d] row-major dense; out may not alias L. Returns 1 if any NaN in out. */
static int SUF(chol_update)(int d, const real *L, const real *x, real coef, real *out, real *w, real *Dv) {
  real b = 1;
  for (int j = 0; j < d; ++j) {
    w[j] = x[j];
    real dg = L[j * d + j];
    Dv[j] = dg * dg;
  }
  /* out <- L / diag (unit-diagonal columns) */
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) out[i * d + j] = L[i * d + j] / L[j * d + j];
  for (int j = 0; j < d; ++j) {
    real wj = w[j];
    real gamma = b * Dv[j] + coef * wj * wj;
    real Dj_new = gamma / b;
    b = gamma / Dv[j];
    real cf = coef * wj / gamma;
    for (int i = 0; i < d; ++i) {
      w[i] = w[i] - wj * out[i * d + j];
      out[i * d + j] = out[i * d + j] + cf * w[i];
    }
    Dv[j] = Dj_new;
  }
  int bad = 0;
  for (int j = 0; j < d; ++j) Dv[j] = R_SQRT(Dv[j]);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      real v = out[i * d + j] * Dv[j];
      out[i * d + j] = v;
      bad |= (v != v);
    }
  return bad;
}

/* ---- Philox4x32-10 + draw mapping (mirrors adaptive_mcmc_b200/csrc/rng.cuh) ---- */
static inline void SUF(philox)(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

static void SUF(philox_words)(uint64_t seed, uint64_t chain, uint64_t step, int n_words, uint32_t *out) {
  int nblk = (n_words + 3) / 4;
  for (int blk = 0; blk < nblk; ++blk) {
    uint32_t c[4] = {(uint32_t)step, (uint32_t)(((step >> 32) & 0xFFFFFFu) | ((uint32_t)blk << 24)),
                     (uint32_t)chain, (uint32_t)(chain >> 32)};
    SUF(philox)(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    for (int k = 0; k < 4; ++k) out[4 * blk + k] = c[k];
  }
}

static void SUF(draws)(uint64_t seed, uint64_t chain, uint64_t step, int d, real *z, real *u, uint32_t *wbuf) {
  int npair = (d + 1) / 2;
  SUF(philox_words)(seed, chain, step, 2 * npair + 1, wbuf);
  for (int p = 0; p < npair; ++p) {
    float u1 = (float)wbuf[2 * p] * 0x1p-32f + 0x1p-33f;
    float th = 6.283185307179586f * ((float)(int32_t)wbuf[2 * p + 1] * 0x1p-32f);
    float r = sqrtf(-2.0f * logf(u1));
    z[2 * p] = (real)(r * cosf(th));
    if (2 * p + 1 < d) z[2 * p + 1] = (real)(r * sinf(th));
  }
  *u = (real)((float)(wbuf[2 * npair] >> 8) * 0x1p-24f);
}

/* q0 ~ U(-radius, radius)^d at the reserved step index 2^56-1 (python/kernels/arwmh.py:111-115) */
void SUF(oracle_init_uniform)(uint64_t seed, int64_t chain_offset, int64_t C, int d, double radius, real *q0) {
  int nw = ((d + 3) / 4) * 4;
  uint32_t *wbuf = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nw);
  for (int64_t c = 0; c < C; ++c) {
    SUF(philox_words)(seed, (uint64_t)(c + chain_offset), ((uint64_t)1 << 56) - 1, d, wbuf);
    for (int k = 0; k < d; ++k) {
      float u = (float)(wbuf[k] >> 8) * 0x1p-24f;
      q0[c * d + k] = (real)((u * 2.0f - 1.0f) * (float)radius);
    }
  }
  free(wbuf);
}

/* ---- model construction helper ---------------------------------------- */
static real *SUF(to_real)(const double *src, int64_t n) {
  real *p = (real *)malloc(sizeof(real) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) p[i] = (real)src[i];
  return p;
}

/* arrays are float64 on input: eight_schools (y[8], sigma[8]); kidiq (kid, hs, iq)[n];
 * diamonds (Xc[n*Kc] centred, Y[n]); gaussian (P[d*d]); std_normal (). */
void SUF(oracle_potential)(int model, int d, int64_t n, const double *a0, const double *a1, const double *a2,
                           int64_t n0, int64_t n1, int64_t n2, int64_t C, const real *q, real *out) {
  SUF(model_t) m = {model, d, n, SUF(to_real)(a0, n0), SUF(to_real)(a1, n1), SUF(to_real)(a2, n2)};
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < C; ++c) out[c] = SUF(potential)(&m, q + c * d);
  free((void *)m.a0); free((void *)m.a1); free((void *)m.a2);
}

/* ---- the run loop: python/kernels/arwmh.py:140-207 per chain ------------ */
/* State arrays are chain-major: z[C,d], loc[C,d], scale[C,d,d] dense row-major.
 * draws: normals[T,C,d], uniforms[T,C] or NULL (Philox).  out_z[S,C,d], out_pe[S,C],
 * out_acc[T,C] (uint8) may be NULL.  Returns 0. */
int SUF(oracle_arwmh_run)(int model, int d, int64_t n_rows, const double *a0, const double *a1, const double *a2,
                          int64_t n0, int64_t n1, int64_t n2, int64_t C, real *z, real *pe, real *macc, real *loc,
                          real *scale, real *lam, real *as_change, int64_t i0, int64_t n_steps,
                          const real *normals, const real *uniforms, uint64_t seed, int64_t chain_offset,
                          int64_t thinning, int64_t collect_start, real *out_z, real *out_pe, uint8_t *out_acc,
                          int64_t num_warmup, double lr_decay, double target, double eps, int adapt,
                          int n_threads) {
  SUF(model_t) m = {model, d, n_rows, SUF(to_real)(a0, n0), SUF(to_real)(a1, n1), SUF(to_real)(a2, n2)};
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    real *nz = (real *)malloc(sizeof(real) * (size_t)(d + 1));
    real *zp = (real *)malloc(sizeof(real) * (size_t)d);
    real *delta = (real *)malloc(sizeof(real) * (size_t)d);
    real *w = (real *)malloc(sizeof(real) * (size_t)d);
    real *Dv = (real *)malloc(sizeof(real) * (size_t)d);
    real *Ls = (real *)malloc(sizeof(real) * (size_t)d * d);
    real *Ln = (real *)malloc(sizeof(real) * (size_t)d * d);
    uint32_t *wbuf = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(d + 8));
#pragma omp for schedule(static)
    for (int64_t c = 0; c < C; ++c) {
      real *zc = z + c * d, *mu = loc + c * d, *L = scale + c * (int64_t)d * d;
      real U = pe[c], ma = macc[c], lm = lam[c], asc = as_change[c];
      for (int64_t t = 0; t < n_steps; ++t) {
        int64_t i = i0 + t;
        real u;
        if (normals) {
          memcpy(nz, normals + (t * C + c) * d, sizeof(real) * (size_t)d);
          u = uniforms[t * C + c];
        } else {
          SUF(draws)(seed, (uint64_t)(c + chain_offset), (uint64_t)i, d, nz, &u, wbuf);
        }
        /* :166-167 */
        real el = R_EXP(lm);
        for (int r = 0; r < d; ++r) {
          real acc = 0;
          for (int k = 0; k < d; ++k) {
            real ps = L[r * d + k] * el + (r == k ? (real)eps : (real)0);
            acc += ps * nz[k];
          }
          zp[r] = zc[r] + acc;
        }
        /* :170-174 */
        real Up = SUF(potential)(&m, zp);
        if (Up != Up) Up = (real)INFINITY;
        real e = R_EXP(U - Up);
        real alpha = (e > (real)1) ? (real)1 : e;
        int acc = (u < alpha);
        if (out_acc) out_acc[t * C + c] = (uint8_t)acc;
        if (acc) { memcpy(zc, zp, sizeof(real) * (size_t)d); U = Up; }
        /* :180-185 */
        int64_t itr = i + 1;
        int64_t n = (i < num_warmup) ? itr : itr - num_warmup;
        real gamma = (real)1 / R_POW((real)n, (real)lr_decay);
        ma = ma + (alpha - ma) / (real)n;
        if (adapt) {
          /* :188-191 */
          real sq = R_SQRT((real)1 - gamma);
          for (int k = 0; k < d; ++k) { delta[k] = zc[k] - mu[k]; mu[k] = mu[k] + gamma * delta[k]; }
          for (int k = 0; k < d * d; ++k) Ls[k] = sq * L[k];
          int bad = SUF(chol_update)(d, Ls, delta, gamma, Ln, w, Dv);
          /* :193 */
          real lm_new = lm + gamma * (alpha - (real)target);
          /* :197 */
          real el_new = R_EXP(lm_new), ss = 0;
          for (int k = 0; k < d * d; ++k) {
            real Lk = bad ? L[k] : Ln[k];
            real df = Lk * el_new - L[k] * el;
            ss += df * df;
            L[k] = Lk;
          }
          asc = R_SQRT(ss);
          lm = lm_new;
        }
        int64_t done = t + 1 - collect_start;
        if (done > 0 && done % thinning == 0) {
          int64_t s = done / thinning - 1;
          if (out_z) memcpy(out_z + (s * C + c) * d, zc, sizeof(real) * (size_t)d);
          if (out_pe) out_pe[s * C + c] = U;
        }
      }
      pe[c] = U; macc[c] = ma; lam[c] = lm; as_change[c] = asc;
    }
    free(nz); free(zp); free(delta); free(w); free(Dv); free(Ls); free(Ln); free(wbuf);
  }
  free((void *)m.a0); free((void *)m.a1); free((void *)m.a2);
  return 0;
}
