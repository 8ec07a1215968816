// arwmh_block.cuh -- block-per-chain fused ARWMH kernel: one CTA owns one chain for the whole launch.
//
// Used where one chain's step has enough internal parallelism to feed a CTA:
//   * diamonds (d = 26): the N = 5000-row likelihood is spread over the CTA's threads (this is the
//     exact fp64 / fp32 CUDA-core parity path and the few-chain path of BASELINE.json configs[1];
//     many-chain diamonds runs on the tcgen05 path, diamonds_tc.cu);
//   * correlated Gaussian d = 200 (configs[4]): the 80 KB proposal factor lives in shared memory.
//
// Chain state lives in SHARED MEMORY across the K fused steps (HBM is touched at launch start/end
// and for the thinned sample stream only).  The proposal factor is carried as L = Lt diag(sqrt(Dg))
// with Lt unit lower triangular, stored COLUMN-MAJOR packed so that "thread i <-> row i" accesses
// of a fixed column are contiguous (bank-conflict free).
#pragma once
#include <cooperative_groups.h>
#include "arwmh_small.cuh"

namespace amcmc {

constexpr int kBlockThreads = 256;       // many chains: several CTAs per SM
constexpr int kBlockThreadsWide = 512;   // few chains (at most one per SM): twice the loads in flight, still 128 registers per thread
constexpr int kWideMaxChainsPerSm = 1;

AMCMC_HD int colbase(int j, int d) { return j * (d - 1) - (j * (j - 1)) / 2; }
// strictly-lower element (i > j) of the column-major packed unit factor
AMCMC_HD int cm_idx(int i, int j, int d) { return colbase(j, d) + (i - j - 1); }

#ifdef __CUDACC__

template <typename R> __device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA; every thread receives the result.  `red` holds NT/32 values.
template <typename R, int NT> __device__ __forceinline__ R block_sum(R v, R* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  R t = (lane < NT / 32) ? red[lane] : (R)0;
  return warp_sum(t);
}

// ---------------------------------------------------------------------------------------------
// diamonds: python/scripts/run_diamonds_lr_decay.py:24-40.  q = [Intercept, b[0..Kc), log sigma]
//   U = 1/2 sum b^2 + 2 log1p(((I-8)/10)^2/3) + 2 log1p((e^s/10)^2/3) - s + N s + 1/2 e^{-2s} RSS + cst
//   RSS = sum_n (Y_n - I - Xc[n,:] b)^2,  Xc = X[:,1:] - column means (precomputed on the host in fp64)
// ---------------------------------------------------------------------------------------------
template <typename R> struct DiamondsBlockModel {
  int d, kc;
  int n;
  int n_stride;          // leading dimension of XcT
  const R* __restrict__ XcT;  // [kc][n_stride]  (column-major: consecutive threads read consecutive rows)
  const R* __restrict__ Y;    // [n]
  double cst;

  static constexpr bool kRowSplit = true;  // the likelihood is a sum over data rows: two CTAs of a cluster can share it

  // sum of squared residuals over rows [r_begin, r_end), reduced over the CTA (every thread gets the sum)
  template <int NT> __device__ R rss_rows(const R* q, R* red, int r_begin, int r_end) const {
    const R icpt = q[0];
    R ss0 = 0, ss1 = 0;
    int r = r_begin + threadIdx.x;
    for (; r + NT < r_end; r += 2 * NT) {  // two rows in flight per thread
      R m0 = icpt, m1 = icpt;
      for (int k = 0; k < kc; ++k) {
        const R bk = q[1 + k];
        m0 = fma(__ldg(XcT + (size_t)k * n_stride + r), bk, m0);
        m1 = fma(__ldg(XcT + (size_t)k * n_stride + r + NT), bk, m1);
      }
      const R e0 = __ldg(Y + r) - m0, e1 = __ldg(Y + r + NT) - m1;
      ss0 = fma(e0, e0, ss0);
      ss1 = fma(e1, e1, ss1);
    }
    if (r < r_end) {
      R m0 = icpt;
      for (int k = 0; k < kc; ++k) m0 = fma(__ldg(XcT + (size_t)k * n_stride + r), q[1 + k], m0);
      const R e0 = __ldg(Y + r) - m0;
      ss0 = fma(e0, e0, ss0);
    }
    return block_sum<R, NT>(ss0 + ss1, red);
  }

  __device__ R finish(const R* q, R rss) const {
    const R icpt = q[0];
    R sb = 0;
    for (int k = 0; k < kc; ++k) sb = fma(q[1 + k], q[1 + k], sb);
    const R s = q[1 + kc];
    const R ti = (icpt - (R)8) * (R)0.1;
    const R ts = Num<R>::exp(s) * (R)0.1;
    const R third = (R)(1.0 / 3.0);
    // N*s, the normalisation constant and e^{-2s}/2 * RSS are O(1e3..1e4) each and cancel to O(1e3): the
    // handful of scalar operations is done in float64 even for the fp32 instantiation (an SFU exp or an
    // fp32 add at 1e4 would each cost ~1e-3 absolute on U)
    const double inv2var = 0.5 * ::exp(-2.0 * (double)s);
    return (R)((double)((R)0.5 * sb + (R)2 * Num<R>::log1p(ti * ti * third) + (R)2 * Num<R>::log1p(ts * ts * third)) +
               ((double)n - 1.0) * (double)s + cst + inv2var * (double)rss);
  }

  template <int NT> __device__ R potential(const R* q, R* red) const { return finish(q, rss_rows<NT>(q, red, 0, n)); }
};

// ---------------------------------------------------------------------------------------------
// Gaussian N(0, Sigma), Sigma^-1 = P P^T:  U = 1/2 |P^T q|^2   (BASELINE.json configs[4])
// P is row-major dense lower-triangular in global memory (L2/L1 resident, shared by all CTAs).
// ---------------------------------------------------------------------------------------------
template <typename R> struct GaussianBlockModel {
  int d;
  const R* __restrict__ P;  // [d][d] row-major, lower triangular
  template <int NT> __device__ R potential(const R* q, R* red) const {
    R acc = 0;
    for (int j = threadIdx.x; j < d; j += NT) {
      R v = 0;
      for (int i = j; i < d; ++i) v = fma(q[i], __ldg(P + (size_t)i * d + j), v);  // coalesced across j
      acc = fma(v, v, acc);
    }
    return (R)0.5 * block_sum<R, NT>(acc, red);
  }
};

// Shared-memory carve-up for one chain of dimension d
template <typename R> struct BlockSmem {
  R *x, *xp, *mu, *Dg, *z, *y, *w, *coef, *Lt, *red, *scal;
  __device__ BlockSmem(unsigned char* base, int d) {
    R* p = reinterpret_cast<R*>(base);
    const int dp = (d + 3) & ~3;
    x = p; p += dp; xp = p; p += dp; mu = p; p += dp; Dg = p; p += dp;
    z = p; p += dp; y = p; p += dp; w = p; p += dp; coef = p; p += dp;
    red = p; p += 32; scal = p; p += 8;
    Lt = p;
  }
  static size_t bytes(int d) {
    const int dp = (d + 3) & ~3;
    return sizeof(R) * ((size_t)8 * dp + 40 + (size_t)d * (d - 1) / 2 + 4);
  }
};

// Rank-one sweep for d <= 32 executed by warp 0: lane i owns row i of Lt (pulled into registers),
// the scalar recurrence runs redundantly on every lane, w_j travels by shuffle.
// Returns (on every lane of warp 0) the squared Frobenius change when WANT.
template <typename R, bool WANT>
__device__ __forceinline__ R sweep_warp(const BlockSmem<R>& sm, int d, R gamma, R el_old, R el_new) {
  const int i = threadIdx.x;  // lane == row (caller guarantees threadIdx.x < 32)
  R row[31];
#pragma unroll
  for (int j = 0; j < 31; ++j) row[j] = (j < i && i < d) ? sm.Lt[cm_idx(i, j, d)] : (R)0;
  R w = (i < d) ? sm.w[i] : (R)0;
  const R Dmine = (i < d) ? sm.Dg[i] : (R)1;
  R t = (R)1, ss = (R)0, Dnew = Dmine;
  const R omg = (R)1 - gamma;
#pragma unroll
  for (int j = 0; j < 31; ++j) {
    if (j < d) {
      const R wj = __shfl_sync(0xffffffffu, w, j);
      const R Dold = __shfl_sync(0xffffffffu, Dmine, j);
      const R Dj = omg * Dold;
      const R cw = gamma * wj;
      const R g = fma(cw * wj, t, Dj);
      const R tr = t * Num<R>::rcp(g);
      const R coef = cw * tr;
      t = Dj * tr;
      if (i == j) Dnew = g;
      R so = 0, sn = 0;
      if (WANT) {
        so = Num<R>::sqrt(Dold) * el_old;
        sn = Num<R>::sqrt(g) * el_new;
        if (i == j) { const R dd = sn - so; ss = fma(dd, dd, ss); }
      }
      if (i > j && i < d) {
        const R Lo = row[j];
        w = fma(-wj, Lo, w);
        const R Ln = fma(coef, w, Lo);
        row[j] = Ln;
        if (WANT) { const R df = fma(Ln, sn, -(Lo * so)); ss = fma(df, df, ss); }
      }
    }
  }
  // last column (j = d-1 when d == 32 has no sub-diagonal entries; handled above for j <= 30; j = 31:)
  if (d == 32) {
    const R wj = __shfl_sync(0xffffffffu, w, 31);
    const R Dold = __shfl_sync(0xffffffffu, Dmine, 31);
    const R Dj = omg * Dold;
    const R g = fma(gamma * wj * wj, t, Dj);
    if (i == 31) Dnew = g;
    if (WANT && i == 31) { const R dd = Num<R>::sqrt(g) * el_new - Num<R>::sqrt(Dold) * el_old; ss = fma(dd, dd, ss); }
  }
  if (i < d) {
    sm.Dg[i] = Dnew;
#pragma unroll
    for (int j = 0; j < 31; ++j)
      if (j < i) sm.Lt[cm_idx(i, j, d)] = row[j];
  }
  return WANT ? warp_sum(ss) : (R)0;
}

// |L|_F^2 for d <= 32 by warp 0
template <typename R> __device__ __forceinline__ R frob2_warp(const BlockSmem<R>& sm, int d) {
  const int j = threadIdx.x;
  R col = 0;
  if (j < d) {
    col = (R)1;
    for (int i = j + 1; i < d; ++i) { const R v = sm.Lt[cm_idx(i, j, d)]; col = fma(v, v, col); }
    col *= sm.Dg[j];
  }
  return warp_sum(col);
}

// Potential of the CTA's chain with the data rows shared by the CL CTAs of the cluster (CL = 1: the model's own potential).
// `ncall` counts the calls of this CTA: the exchange slots alternate, so one cluster barrier per call is enough.
template <class BM, typename R, int NT, int CL>
__device__ __forceinline__ R cluster_potential(const BM& m, const R* q, R* red, R (*xch)[CL], int cl_rank, unsigned& ncall) {
  if constexpr (CL > 1) {
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const int share = ((m.n + CL - 1) / CL + 31) & ~31, par = (int)(ncall & 1u);
    ++ncall;
    const int r0 = min(m.n, cl_rank * share), r1 = min(m.n, r0 + share);
    const R part = m.template rss_rows<NT>(q, red, r0, r1);
    if (threadIdx.x < CL) *cluster.map_shared_rank(&xch[par][cl_rank], threadIdx.x) = part;  // my partial into every CTA of the cluster
    cluster.sync();
    R rss = xch[par][0];
#pragma unroll
    for (int k = 1; k < CL; ++k) rss += xch[par][k];  // rank order: the same sum in every CTA
    return m.finish(q, rss);
  } else {
    return m.template potential<NT>(q, red);
  }
}

// CL = 2, 4, 8 (few-chain diamonds runs, launched with that cluster dimension): the CTAs of a cluster carry the SAME chain -- identical draws,
// proposals, decisions and adaptation, so no state is exchanged -- and split the data rows of the likelihood; the partial
// sums of squared residuals cross through distributed shared memory (one cluster barrier per step, two slot sets alternate)
// and are added in rank order in every CTA.  Rank 0 writes the outputs.
template <class BM, typename R, bool ADAPT, bool EXTERNAL, int NT, int CL = 1>
__global__ void __launch_bounds__(NT)
arwmh_block_kernel(const BM m, const StateView<R> st, const RunView<R> a, const int d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ R cl_xch[2][CL];
  BlockSmem<R> sm(smem_raw, d);
  const int tid = threadIdx.x;
  const int64_t C = st.C;
  const int64_t c = blockIdx.x / CL;
  const int cl_rank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
  const bool lead = cl_rank == 0;
  unsigned cl_calls = 0;
  if (CL > 1) cooperative_groups::this_cluster().sync();  // the peer's exchange slots exist
  // ---- load the chain: L (row-major packed, with diagonal) -> Lt (column-major packed), Dg
  for (int k = tid; k < d; k += NT) {
    sm.x[k] = st.z[k * C + c];
    sm.mu[k] = st.loc[k * C + c];
    const R dg = st.scale[(int64_t)tri_full(k, k) * C + c];
    sm.Dg[k] = dg * dg;
    sm.y[k] = (R)1 / dg;  // temporarily 1/diag
  }
  __syncthreads();
  for (int e = tid; e < d * (d - 1) / 2; e += NT) {
    // e enumerates strictly-lower (i, j) row-major: i(i-1)/2 + j
    int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);
    while (i * (i - 1) / 2 > e) --i;
    while ((i + 1) * i / 2 <= e) ++i;
    const int j = e - i * (i - 1) / 2;
    sm.Lt[cm_idx(i, j, d)] = st.scale[(int64_t)tri_full(i, j) * C + c] * sm.y[j];
  }
  R U = st.pe[c], lam = st.lam[c], macc = st.macc[c], asc = st.asc[c];
  __syncthreads();

  const Philox rng(a.seed, (uint64_t)(c + a.chain_offset));
  const int npair = (d + 1) / 2;
  int64_t until_collect = a.collect_start + a.thinning;
  int64_t sidx = 0;
  for (int64_t t = 0; t < a.n_steps; ++t) {
    const int64_t it = a.i0 + t;
    // ---- draws: z[0..d), u
    if (EXTERNAL) {
      for (int k = tid; k < d; k += NT) sm.z[k] = a.normals[(t * d + k) * C + c];
      if (tid == 0) sm.scal[0] = a.uniforms[t * C + c];
    } else {
      for (int p = tid; p <= npair; p += NT) {
        uint32_t o[4];
        if (p < npair) {
          rng.block((uint64_t)it, (uint32_t)(p >> 1), o);
          float z0, z1;
          box_muller(o[(p & 1) * 2], o[(p & 1) * 2 + 1], z0, z1);
          sm.z[2 * p] = (R)z0;
          if (2 * p + 1 < d) sm.z[2 * p + 1] = (R)z1;
        } else {
          rng.block((uint64_t)it, (uint32_t)((2 * npair) >> 2), o);
          sm.scal[0] = (R)word_to_uniform(o[(2 * npair) & 3]);
        }
      }
    }
    __syncthreads();
    // ---- proposal  x' = x + e^lam Lt (sqrt(Dg) .* z) + eps z      (arwmh.py:166-167)
    const R el = Num<R>::exp(lam);
    for (int k = tid; k < d; k += NT) sm.y[k] = sm.z[k] * Num<R>::sqrt(sm.Dg[k]);
    __syncthreads();
    for (int i = tid; i < d; i += NT) {
      R acc = sm.y[i];
      for (int j = 0; j < i; ++j) acc = fma(sm.Lt[cm_idx(i, j, d)], sm.y[j], acc);
      sm.xp[i] = sm.x[i] + fma(el, acc, a.eps * sm.z[i]);
    }
    __syncthreads();
    // ---- potential, accept (:170-178); every thread holds the same scalars
    R Up = cluster_potential<BM, R, NT, CL>(m, sm.xp, sm.red, cl_xch, cl_rank, cl_calls);
    if (Num<R>::isnan(Up)) Up = Num<R>::inf();
    const R e = Num<R>::exp(U - Up);
    const R alpha = (e > (R)1) ? (R)1 : e;
    const bool acc = sm.scal[0] < alpha;
    if (acc) {
      for (int k = tid; k < d; k += NT) sm.x[k] = sm.xp[k];
      U = Up;
    }
    if (a.out_acc && tid == 0 && lead) a.out_acc[t * C + c] = (uint8_t)acc;
    const int64_t n = (it < a.num_warmup) ? (it + 1) : (it + 1 - a.num_warmup);
    // :185 running mean over n; the frozen kernel (sample_Pnx, pooled windows) reports the mean over THIS launch
    const R nf = (R)n;
    macc = fma(alpha - macc, Num<R>::rcp(ADAPT ? nf : (R)(t + 1)), macc);
    if (ADAPT) {
      const bool n_is_one = (n == 1);
      const R gamma = n_is_one ? (R)1 : Num<R>::pow_neg(nf, a.lr_decay);
      const bool last = (t == a.n_steps - 1);
      __syncthreads();  // x final
      int ok_local = 1;
      for (int k = tid; k < d; k += NT) {
        const R dl = sm.x[k] - sm.mu[k];
        sm.mu[k] = fma(gamma, dl, sm.mu[k]);
        sm.w[k] = dl;
        ok_local &= (Num<R>::abs(dl) < Num<R>::kBig) && (sm.Dg[k] > (R)0);
      }
      const int ok = __syncthreads_and(ok_local) && !n_is_one;
      const R lam_new = fma(gamma, alpha - a.target, lam);
      const R el_new = Num<R>::exp(lam_new);
      if (tid < 32) {  // d <= 32: warp 0 sweeps
        R ss = 0;
        if (ok) ss = last ? sweep_warp<R, true>(sm, d, gamma, el, el_new) : sweep_warp<R, false>(sm, d, gamma, el, el_new);
        else if (last) { const R de = el_new - el; ss = de * de * frob2_warp(sm, d); }
        if (last && tid == 0) sm.scal[1] = Num<R>::sqrt(ss);
      }
      lam = lam_new;
      __syncthreads();
      if (last) asc = sm.scal[1];
    } else {
      __syncthreads();
    }
    if (--until_collect == 0) {
      until_collect = a.thinning;
      if (a.out_z && lead)
        for (int k = tid; k < d; k += NT) a.out_z[(sidx * d + k) * C + c] = sm.x[k];
      if (a.out_pe && tid == 0 && lead) a.out_pe[sidx * C + c] = U;
      ++sidx;
    }
  }
  if (CL > 1) cooperative_groups::this_cluster().sync();  // nobody leaves while the peer may still write into its slots
  if (!lead) return;
  // ---- store
  __syncthreads();
  for (int k = tid; k < d; k += NT) st.z[k * C + c] = sm.x[k];
  if (tid == 0) st.pe[c] = U;
  if (ADAPT) {
    for (int k = tid; k < d; k += NT) {
      st.loc[k * C + c] = sm.mu[k];
      const R sd = ::sqrt(sm.Dg[k]);
      sm.y[k] = sd;
      st.scale[(int64_t)tri_full(k, k) * C + c] = sd;
    }
    __syncthreads();
    for (int e2 = tid; e2 < d * (d - 1) / 2; e2 += NT) {
      int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e2)) * 0.5f);
      while (i * (i - 1) / 2 > e2) --i;
      while ((i + 1) * i / 2 <= e2) ++i;
      const int j = e2 - i * (i - 1) / 2;
      st.scale[(int64_t)tri_full(i, j) * C + c] = sm.Lt[cm_idx(i, j, d)] * sm.y[j];
    }
    if (tid == 0) { st.lam[c] = lam; st.asc[c] = asc; }
  }
  if (tid == 0) st.macc[c] = macc;
}

// ARWMH.init for block models: q0 (given or U(-r,r)), U0, loc = q0, scale = I, ...
template <class BM, typename R>
__global__ void __launch_bounds__(kBlockThreads)
arwmh_block_init_kernel(const BM m, const StateView<R> st, const int d, uint64_t seed, int64_t chain_offset, R radius,
                        int use_given_z) {
  constexpr int NT = kBlockThreads;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* q = reinterpret_cast<R*>(smem_raw);
  R* red = q + ((d + 3) & ~3);
  const int tid = threadIdx.x;
  const int64_t C = st.C, c = blockIdx.x;
  const Philox rng(seed, (uint64_t)(c + chain_offset));
  for (int k = tid; k < d; k += NT) {
    R v;
    if (use_given_z) v = st.z[k * C + c];
    else {
      uint32_t o[4];
      rng.block(kInitStep, (uint32_t)(k >> 2), o);
      v = (R)((word_to_uniform(o[k & 3]) * 2.0f - 1.0f) * (float)radius);
      st.z[k * C + c] = v;
    }
    q[k] = v;
    st.loc[k * C + c] = v;
  }
  __syncthreads();
  const R U = m.template potential<NT>(q, red);
  for (int e = tid; e < d * (d + 1) / 2; e += NT) {
    int i = (int)((sqrtf(1.0f + 8.0f * (float)e) - 1.0f) * 0.5f);
    while (i * (i + 1) / 2 > e) --i;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    const int j = e - i * (i + 1) / 2;
    st.scale[(int64_t)e * C + c] = (i == j) ? (R)1 : (R)0;
  }
  if (tid == 0) { st.pe[c] = U; st.lam[c] = 0; st.macc[c] = 0; st.asc[c] = 0; }
}

template <class BM, typename R>
__global__ void __launch_bounds__(kBlockThreads)
potential_block_kernel(const BM m, const int d, int64_t n, const R* __restrict__ qs, R* __restrict__ out) {
  constexpr int NT = kBlockThreads;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* q = reinterpret_cast<R*>(smem_raw);
  R* red = q + ((d + 3) & ~3);
  const int64_t c = blockIdx.x;
  for (int k = threadIdx.x; k < d; k += NT) q[k] = qs[k * n + c];
  __syncthreads();
  const R U = m.template potential<NT>(q, red);
  if (threadIdx.x == 0) out[c] = U;
}

#endif  // __CUDACC__
}  // namespace amcmc
