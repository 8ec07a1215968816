// asss_small.cuh -- adaptive stereographic slice sampler (python/kernels/asss.py:192-269), thread-per-chain
// with the whole chain in registers, for the same small-d models as arwmh_small.cuh.  SURVEY 8f rank 2.
//
// One step: project x to the sphere S^d with the adapted (loc, scale*sqrt(d)) (asss.py:33-44, :218, :227);
// draw a tangent direction v and a level t = pe(z) - log u (:231-237); shrink a great-circle bracket until
// the transformed energy pe(z cos th + v sin th) = U(x(z)) + d log(1 - z_{d+1}) drops below t, at most 50
// times (:59-96); map back (:241); adapt loc / scale exactly like ARWMH (:246-255, no step size).
// `as_change` = |loc' - loc| + |scale' - scale|_F (:259-267).  The proposal factor is carried as
// L = Lt diag(sqrt(Dg)) like in arwmh_small.cuh.  mean_accept_prob is reused to carry the running mean number
// of shrinkage iterations (the reference state has no such field).
#pragma once
#include "arwmh_small.cuh"

namespace amcmc {

constexpr int kAsssMaxIter = 50;
constexpr int kAsssUniforms = 2 + kAsssMaxIter;  // external uniforms per step: u_t, theta_0/2pi, 50 shrink draws

template <typename R> struct SinCos;
template <> struct SinCos<float> {
  static AMCMC_HD void eval(float t, float& s, float& c) {
#ifdef __CUDA_ARCH__
    __sincosf(t, &s, &c);
#else
    s = ::sinf(t); c = ::cosf(t);
#endif
  }
};
template <> struct SinCos<double> {
  static AMCMC_HD void eval(double t, double& s, double& c) { s = ::sin(t); c = ::cos(t); }
};

// x(z) = loc + (L + eps I) sqrt(d) * z_{1:d} / (1 - z_{d+1})   and the transformed energy
template <class Model, typename R>
AMCMC_HD R asss_transformed(const ChainRegs<R, Model::D>& s, const Model& m, const R (&cs)[Model::D], R eps_dsq,
                            const R (&zc)[Model::D], R zl, R (&xn)[Model::D], R& Un, R& om) {
  constexpr int D = Model::D;
  om = (R)1 - zl;
  const R rb = (R)1 / om;
  R xb[D];
#pragma unroll
  for (int k = 0; k < D; ++k) xb[k] = zc[k] * rb;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    R acc = fma(cs[i] + eps_dsq, xb[i], s.mu[i]);
#pragma unroll
    for (int j = 0; j < i; ++j) acc = fma(s.Lt[tri_strict(i, j)] * cs[j], xb[j], acc);
    xn[i] = acc;
  }
  Un = m.potential(xn);
  R pe = Un + (R)D * Num<R>::log(om);
  if (Num<R>::isnan(pe)) pe = Num<R>::inf();
  return pe;
}

// vn: D+1 normals; u_t, u_th: uniforms; next_u(k): k-th shrinkage uniform (lazily evaluated)
template <class Model, typename R, bool ADAPT, class NextU>
AMCMC_HD int asss_step(ChainRegs<R, Model::D>& s, const Model& m, const R (&vn)[Model::D + 1], R u_t, R u_th,
                       NextU next_u, R nf, bool n_is_one, R lr_decay, R eps, bool want_asc) {
  constexpr int D = Model::D;
  const R dsq = Num<R>::sqrt((R)D);
  const R eps_dsq = eps * dsq;
  R cs[D];  // column scale sqrt(Dg_j) * sqrt(d)
#pragma unroll
  for (int j = 0; j < D; ++j) cs[j] = Num<R>::sqrt(s.Dg[j]) * dsq;
  // ---- project (asss.py:33-44): y = ((L + eps I) sqrt(d))^-1 (x - loc)
  R y[D], nsq = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    R acc = s.x[i] - s.mu[i];
#pragma unroll
    for (int j = 0; j < i; ++j) acc = fma(-s.Lt[tri_strict(i, j)] * cs[j], y[j], acc);
    y[i] = acc / (cs[i] + eps_dsq);
    nsq = fma(y[i], y[i], nsq);
  }
  R zc[D], zl;
  {
    const R r = (R)1 / (nsq + (R)1);
#pragma unroll
    for (int k = 0; k < D; ++k) zc[k] = (R)2 * y[k] * r;
    zl = (nsq - (R)1) * r;
  }
  const R pe_z = s.U + (R)D * Num<R>::log((R)1 - zl);  // :228 (stored U == U(x(z)) up to round-off)
  // ---- tangent direction (:231-233)
  R vc[D], vl = vn[D];
  R dot = vl * zl;
#pragma unroll
  for (int k = 0; k < D; ++k) { vc[k] = vn[k]; dot = fma(vc[k], zc[k], dot); }
  R vsq = 0;
#pragma unroll
  for (int k = 0; k < D; ++k) { vc[k] = fma(-dot, zc[k], vc[k]); vsq = fma(vc[k], vc[k], vsq); }
  vl = fma(-dot, zl, vl);
  vsq = fma(vl, vl, vsq);
  {
    const R rn = (R)1 / Num<R>::sqrt(vsq);
#pragma unroll
    for (int k = 0; k < D; ++k) vc[k] *= rn;
    vl *= rn;
  }
  const R t_pe = pe_z - Num<R>::log(u_t);  // :236-237
  // ---- shrinkage (:59-96)
  const R two_pi = (R)6.283185307179586476925;
  R theta = two_pi * u_th, th_min = theta - two_pi, th_max = theta;
  int iter = 0;
  R xn[D], Un = 0;
  while (true) {
    R sn, cn;
    SinCos<R>::eval(theta, sn, cn);
    R ztc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) ztc[k] = fma(zc[k], cn, vc[k] * sn);
    const R ztl = fma(zl, cn, vl * sn);
    R om;
    const R pe = asss_transformed<Model, R>(s, m, cs, eps_dsq, ztc, ztl, xn, Un, om);
    const bool cont = (iter < kAsssMaxIter) && ((pe > t_pe) || (om < eps));
    if (!cont) break;
    if (theta < (R)0) th_min = theta; else th_max = theta;
    theta = fma(next_u(iter), th_max - th_min, th_min);
    ++iter;
  }
  if (iter >= kAsssMaxIter) {  // :94  give up: theta = 0 (stay at z)
    R om;
    asss_transformed<Model, R>(s, m, cs, eps_dsq, zc, zl, xn, Un, om);
  }
  if (Num<R>::isnan(Un)) Un = Num<R>::inf();  // :244
  if (!ADAPT) {  // sample_Pnx (:279-315): the step is taken with the given adapt_state, no update
#pragma unroll
    for (int k = 0; k < D; ++k) s.x[k] = xn[k];
    s.U = Un;
    return iter;
  }
  // ---- adaptation (:246-267)
  const R gamma = n_is_one ? (R)1 : Num<R>::pow_neg(nf, lr_decay);
  R w[D], dn = 0;
  bool ok = !n_is_one;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    s.x[k] = xn[k];
    const R dl = xn[k] - s.mu[k];
    s.mu[k] = fma(gamma, dl, s.mu[k]);
    w[k] = dl;
    dn = fma(dl, dl, dn);
    ok = ok && (Num<R>::abs(dl) < Num<R>::kBig) && (s.Dg[k] > (R)0);
  }
  s.U = Un;
  if (want_asc) {
    R ss = 0;
    if (ok) ss = rank1_sweep<R, D, true>(s, w, gamma, (R)1, (R)1);
    s.asc = gamma * Num<R>::sqrt(dn) + Num<R>::sqrt(ss);
  } else if (ok) {
    rank1_sweep<R, D, false>(s, w, gamma, (R)1, (R)1);
  }
  s.macc = fma((R)iter - s.macc, Num<R>::rcp(nf), s.macc);  // running mean of shrinkage iterations
  return iter;
}

// Philox draws of one ASSS step: D+1 normals from word pairs, u_t and theta_0 from the next two words;
// the k-th shrinkage uniform is word k%4 of block 64 + k/4 (evaluated lazily, four at a time).
template <typename R, int D>
AMCMC_HD void asss_philox_head(const Philox& g, uint64_t step, R (&vn)[D + 1], R& u_t, R& u_th) {
  constexpr int NPAIR = (D + 2) / 2;
  constexpr int NW = 2 * NPAIR + 2;
  constexpr int NBLK = (NW + 3) / 4;
  uint32_t w[NBLK * 4];
#pragma unroll
  for (int b = 0; b < NBLK; ++b) {
    uint32_t o[4];
    g.block(step, (uint32_t)b, o);
    w[4 * b] = o[0]; w[4 * b + 1] = o[1]; w[4 * b + 2] = o[2]; w[4 * b + 3] = o[3];
  }
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) {
    float a, b;
    box_muller(w[2 * p], w[2 * p + 1], a, b);
    vn[2 * p] = (R)a;
    if (2 * p + 1 < D + 1) vn[2 * p + 1] = (R)b;
  }
  u_t = (R)word_to_uniform(w[2 * NPAIR]);
  u_th = (R)word_to_uniform(w[2 * NPAIR + 1]);
}

// Steps [t0, t1) of the launch for one chain whose state is in `s`; any split of [0, n_steps) into consecutive ranges gives
// the same trajectory and samples (the balanced launch of arwmh_small.cuh hands chains from warp to warp between ranges).
template <class Model, typename R, bool EXTERNAL, bool ADAPT, bool RING = false>
AMCMC_HD void asss_chain_range(ChainRegs<R, Model::D>& s, const Model& m, const RunView<R>& a, const Philox& rng, int64_t C,
                               int64_t c, int64_t t0, int64_t t1) {
  constexpr int D = Model::D;
  int64_t sidx = t0 > a.collect_start ? (t0 - a.collect_start) / a.thinning : 0;  // samples taken before this range
  int64_t until_collect = a.collect_start + (sidx + 1) * a.thinning - t0;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t i = a.i0 + t;
    R vn[D + 1], u_t, u_th;
    if (RING) {  // asss_small_duo_kernel: the head draws of the step come from the CTA's producer warp
#ifdef __CUDA_ARCH__
      __syncthreads();
      const R* buf = a.ring + (size_t)(t & 1) * (D + 3) * 32 + (threadIdx.x & 31);
#pragma unroll
      for (int k = 0; k < D + 1; ++k) vn[k] = buf[k * 32];
      u_t = buf[(D + 1) * 32];
      u_th = buf[(D + 2) * 32];
#endif
    } else if (EXTERNAL) {
#pragma unroll
      for (int k = 0; k < D + 1; ++k) vn[k] = a.normals[(t * (D + 1) + k) * C + c];
      u_t = a.uniforms[(t * kAsssUniforms + 0) * C + c];
      u_th = a.uniforms[(t * kAsssUniforms + 1) * C + c];
    } else {
      asss_philox_head<R, D>(rng, (uint64_t)i, vn, u_t, u_th);
    }
    uint32_t cache[4];
    int cached_blk = -1;
    auto next_u = [&](int k) -> R {
      if (EXTERNAL) return a.uniforms[(t * kAsssUniforms + 2 + k) * C + c];
      const int blk = 64 + (k >> 2);
      if (blk != cached_blk) { rng.block((uint64_t)i, (uint32_t)blk, cache); cached_blk = blk; }
      return (R)word_to_uniform(cache[k & 3]);
    };
    const int64_t n = (i < a.num_warmup) ? (i + 1) : (i + 1 - a.num_warmup);
    const bool last = (t == a.n_steps - 1);
    asss_step<Model, R, ADAPT>(s, m, vn, u_t, u_th, next_u, (R)n, n == 1, a.lr_decay, a.eps, last);
    if (--until_collect == 0) {
      until_collect = a.thinning;
      if (a.out_z) {
#pragma unroll
        for (int k = 0; k < D; ++k) a.out_z[(sidx * D + k) * C + c] = s.x[k];
      }
      if (a.out_pe) a.out_pe[sidx * C + c] = s.U;
      ++sidx;
    }
  }
}

template <class Model, typename R, bool EXTERNAL, bool ADAPT>
AMCMC_HD void asss_chain_run(const Model& m, const StateView<R>& st, const RunView<R>& a, int64_t c) {
  constexpr int D = Model::D;
  ChainRegs<R, D> s;
  load_chain(s, st, c);
  const Philox rng(a.seed, (uint64_t)(c + a.chain_offset));
  asss_chain_range<Model, R, EXTERNAL, ADAPT>(s, m, a, rng, st.C, c, 0, a.n_steps);
  store_chain<R, D, ADAPT>(s, st, c);
}

// range runner for arwmh_small_balanced_kernel
template <class Model, typename R, bool ADAPT, bool EXTERNAL> struct AsssRange {
  static AMCMC_HD void run(ChainRegs<R, Model::D>& s, const Model& m, const RunView<R>& a, const Philox& rng, int64_t C, int64_t c,
                           int64_t t0, int64_t t1) {
    asss_chain_range<Model, R, EXTERNAL, ADAPT>(s, m, a, rng, C, c, t0, t1);
  }
};

#ifdef __CUDACC__
// Few chains: a producer warp generates the head draws (d + 1 normals, u_t, theta_0) of step t + 1 while the chains' warp runs
// step t (see arwmh_small_duo_kernel); the shrink uniforms stay lazy in the chains' warp.  Every output equal to
// asss_small_kernel's.
template <class Model, typename R, bool ADAPT>
__global__ void __launch_bounds__(64) asss_small_duo_kernel(const Model m, const StateView<R> st, const RunView<R> a) {
  constexpr int D = Model::D;
  __shared__ R ring[2 * (D + 3) * 32];
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const int64_t cc = c < st.C ? c : st.C - 1;
  const Philox rng(a.seed, (uint64_t)(cc + a.chain_offset));
  if (threadIdx.x >= 32) {
    for (int64_t t = 0; t < a.n_steps; ++t) {
      R vn[D + 1], u_t, u_th;
      asss_philox_head<R, D>(rng, (uint64_t)(a.i0 + t), vn, u_t, u_th);
      R* buf = ring + (size_t)(t & 1) * (D + 3) * 32 + lane;
#pragma unroll
      for (int k = 0; k < D + 1; ++k) buf[k * 32] = vn[k];
      buf[(D + 1) * 32] = u_t;
      buf[(D + 2) * 32] = u_th;
      __syncthreads();
    }
    return;
  }
  RunView<R> a2 = a;
  a2.ring = ring;
  ChainRegs<R, D> s;
  load_chain(s, st, cc);
  asss_chain_range<Model, R, false, ADAPT, true>(s, m, a2, rng, st.C, cc, 0, a.n_steps);
  if (c < st.C) store_chain<R, D, ADAPT>(s, st, c);
}

template <class Model, typename R, bool EXTERNAL, bool ADAPT>
// 7 resident CTAs of 64 threads per SM (128 registers): 65,536 chains fit in ONE wave, as for arwmh_small_kernel
__global__ void __launch_bounds__(64, (sizeof(R) == 4 ? 7 : 1))
asss_small_kernel(const Model m, const StateView<R> st, const RunView<R> a) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= st.C) return;
  asss_chain_run<Model, R, EXTERNAL, ADAPT>(m, st, a, c);
}
#endif

}  // namespace amcmc
