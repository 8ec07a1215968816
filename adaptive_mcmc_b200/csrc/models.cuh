// models.cuh -- model log-densities (potential energy in unconstrained space) as functors that the
// fused sampler kernels inline.  Each is the closed form of the NumPyro model the reference scripts
// define; flat parameter order = ravel_pytree order of the unconstrained site dict (sorted names).
#pragma once
#include "common.cuh"

namespace amcmc {

constexpr double kLog2PiHalf = 0.91893853320467274178;

// potential_fn = 0.5*|x|^2 : python/jupyter/asumptions_check.ipynb cells 17-28 (N(0,I_d) target)
template <typename R, int D_> struct StdNormalModel {
  static constexpr int D = D_;
  AMCMC_HD R potential(const R (&q)[D]) const {
    R s = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) s = fma(q[k], q[k], s);
    return (R)0.5 * s;
  }
};

// python/scripts/run_eight_schools_lr_decay.py:26-35 -- non-centred eight schools.
// q = [mu, log tau, theta_base[0..7]].
//   mu ~ N(0,5); tau ~ HalfCauchy(5) (+ log-Jacobian t); theta_base ~ N(0,1);
//   y_j ~ N(mu + tau*theta_base_j, sigma_j)
// All additive constants are folded into `cst` on the host (in float64).
template <typename R> struct EightSchoolsModel {
  static constexpr int D = 10;
  R y[8];
  R inv_sigma[8];
  R cst;
  AMCMC_HD R potential(const R (&q)[10]) const {
    const R mu = q[0], t = q[1];
    const R tau = Num<R>::exp(t);
    const R a = mu * (R)0.2;
    const R tq = tau * (R)0.2;
    R U = (R)0.5 * a * a + Num<R>::log1p(tq * tq) - t;
    R se = 0, sr = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const R e = q[2 + j];
      se = fma(e, e, se);
      const R r = (y[j] - fma(tau, e, mu)) * inv_sigma[j];
      sr = fma(r, r, sr);
    }
    return U + (R)0.5 * (se + sr) + cst;
  }
};

// python/scripts/run_kidiq_kidscore_lr_decay.py:29-41 -- q = [beta0, beta1, beta2, log sigma];
// flat prior on beta; sigma ~ HalfCauchy(2.5); kid ~ N(b0 + b1*hs + b2*iq, sigma).
// Data rows are read through the read-only path (every lane reads the same address: broadcast).
template <typename R> struct KidiqModel {
  static constexpr int D = 4;
  const R* __restrict__ kid;
  const R* __restrict__ hs;
  const R* __restrict__ iq;
  int n;
  R cst;  // -(log2 - log pi - log 2.5) + n*0.5*log(2 pi)
  AMCMC_HD R potential(const R (&q)[4]) const {
    const R b0 = q[0], b1 = q[1], b2 = q[2], s = q[3];
    const R sig = Num<R>::exp(s);
    const R sq = sig * (R)0.4;
    R ss = 0;
    for (int i = 0; i < n; ++i) {
#ifdef __CUDA_ARCH__
      const R r = __ldg(kid + i) - fma(b2, __ldg(iq + i), fma(b1, __ldg(hs + i), b0));
#else
      const R r = kid[i] - fma(b2, iq[i], fma(b1, hs[i], b0));
#endif
      ss = fma(r, r, ss);
    }
    const R inv_var = Num<R>::exp((R)-2 * s);
    return Num<R>::log1p(sq * sq) - s + (R)n * s + (R)0.5 * inv_var * ss + cst;
  }
};

}  // namespace amcmc
