// asss_block.cuh -- adaptive stereographic slice sampler (python/kernels/asss.py:33-96, 192-269), block-per-chain:
// one CTA owns one chain, the model's potential is evaluated cooperatively by the CTA (diamonds: 5000 rows over 256
// threads), the d <= 32 vector algebra of the step runs in warp 0 with lane i <-> coordinate i.  The reference runs
// ASSS on diamonds too (run_diamonds_wasserstein.py 'sss'; 3671 it/s recorded, posteriordb_diamonds.ipynb:L1992).
//
// Same arithmetic, draw layout and state conventions as asss_small.cuh (which serves the register-sized models):
//   project x to S^d with (loc, (L + eps I) sqrt(d)); tangent direction from d+1 normals; level t = pe(z) - log u;
//   shrink a great-circle bracket until pe(z cos th + v sin th) <= t, at most 50 times; map back; adapt loc and the
//   LDL^T factor like ARWMH (no step size).  mean_accept_prob carries the running mean number of shrink iterations.
#pragma once
#include "arwmh_block.cuh"
#include "asss_small.cuh"

namespace amcmc {

#ifdef __CUDACC__

// CL > 1: the CTAs of a thread-block cluster carry the same chain and share the data rows of every potential evaluation
// (cluster_potential, arwmh_block.cuh); the shrink loop runs the same number of times in every CTA because all of them see the
// same energies.
template <class BM, typename R, bool EXTERNAL, bool ADAPT, int NT, int CL = 1>
__global__ void __launch_bounds__(NT)
asss_block_kernel(const BM m, const StateView<R> st, const RunView<R> a, const int d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ R cl_xch[2][CL];
  BlockSmem<R> sm(smem_raw, d);
  R* zc = sm.z;      // point on the sphere, first d coordinates
  R* vc = sm.y;      // tangent direction, first d coordinates
  R* cs = sm.coef;   // column scale sqrt(Dg_j) * sqrt(d)
  R* xb = sm.w;      // scratch: normals of the step, then x_base of the current angle, then delta
  const int tid = threadIdx.x;
  const int64_t C = st.C;
  const int64_t c = blockIdx.x / CL;
  const int cl_rank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
  const bool lead = cl_rank == 0;
  unsigned cl_calls = 0;
  if (CL > 1) cooperative_groups::this_cluster().sync();  // the peers' exchange slots exist
  // ---- load the chain (as arwmh_block_kernel)
  for (int k = tid; k < d; k += NT) {
    sm.x[k] = st.z[k * C + c];
    sm.mu[k] = st.loc[k * C + c];
    const R dg = st.scale[(int64_t)tri_full(k, k) * C + c];
    sm.Dg[k] = dg * dg;
    sm.y[k] = (R)1 / dg;
  }
  __syncthreads();
  for (int e = tid; e < d * (d - 1) / 2; e += NT) {
    int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);
    while (i * (i - 1) / 2 > e) --i;
    while ((i + 1) * i / 2 <= e) ++i;
    const int j = e - i * (i - 1) / 2;
    sm.Lt[cm_idx(i, j, d)] = st.scale[(int64_t)tri_full(i, j) * C + c] * sm.y[j];
  }
  R U = st.pe[c], macc = st.macc[c], asc = st.asc[c];
  __syncthreads();

  const Philox rng(a.seed, (uint64_t)(c + a.chain_offset));
  const int npair = (d + 2) / 2;  // pairs covering the d + 1 normals
  const R dsq = Num<R>::sqrt((R)d);
  const R eps_dsq = a.eps * dsq;
  const R two_pi = (R)6.283185307179586476925;
  int64_t until_collect = a.collect_start + a.thinning;
  int64_t sidx = 0;

  // x(z) = loc + (L + eps I) sqrt(d) z_{1:d} / (1 - z_{d+1}) and its energy; every thread returns the same values
  auto transformed = [&](R cn, R sn, R& Un, R& om) -> R {
    const R ztl = fma(sm.scal[0], cn, sm.scal[1] * sn);
    om = (R)1 - ztl;
    const R rb = (R)1 / om;
    __syncthreads();  // xb / xp free
    for (int k = tid; k < d; k += NT) xb[k] = fma(zc[k], cn, vc[k] * sn) * rb;
    __syncthreads();
    for (int i = tid; i < d; i += NT) {
      R acc = fma(cs[i] + eps_dsq, xb[i], sm.mu[i]);
      for (int j = 0; j < i; ++j) acc = fma(sm.Lt[cm_idx(i, j, d)] * cs[j], xb[j], acc);
      sm.xp[i] = acc;
    }
    __syncthreads();
    Un = cluster_potential<BM, R, NT, CL>(m, sm.xp, sm.red, cl_xch, cl_rank, cl_calls);
    R pe = Un + (R)d * Num<R>::log(om);
    if (Num<R>::isnan(pe)) pe = Num<R>::inf();
    return pe;
  };

  for (int64_t t = 0; t < a.n_steps; ++t) {
    const int64_t it = a.i0 + t;
    // ---- draws: d + 1 normals -> xb[0..d), scal[5]; u_t -> scal[6]; theta_0 / 2 pi -> scal[7]
    if (EXTERNAL) {
      for (int k = tid; k <= d; k += NT) {
        const R v = a.normals[(t * (d + 1) + k) * C + c];
        if (k < d) xb[k] = v; else sm.scal[5] = v;
      }
      if (tid == 0) {
        sm.scal[6] = a.uniforms[(t * kAsssUniforms + 0) * C + c];
        sm.scal[7] = a.uniforms[(t * kAsssUniforms + 1) * C + c];
      }
    } else {
      for (int p = tid; p <= npair; p += NT) {
        uint32_t o[4];
        if (p < npair) {
          rng.block((uint64_t)it, (uint32_t)(p >> 1), o);
          float z0, z1;
          box_muller(o[(p & 1) * 2], o[(p & 1) * 2 + 1], z0, z1);
          if (2 * p < d) xb[2 * p] = (R)z0; else sm.scal[5] = (R)z0;
          if (2 * p + 1 < d) xb[2 * p + 1] = (R)z1; else if (2 * p + 1 == d) sm.scal[5] = (R)z1;
        } else {
          rng.block((uint64_t)it, (uint32_t)((2 * npair) >> 2), o);
          sm.scal[6] = (R)word_to_uniform(o[(2 * npair) & 3]);
          sm.scal[7] = (R)word_to_uniform(o[(2 * npair + 1) & 3]);
        }
      }
    }
    __syncthreads();
    // ---- warp 0: projection (asss.py:33-44, :218, :227), tangent direction (:231-233), level (:236-237)
    if (tid < 32) {
      const int i = tid;
      const bool on = i < d;
      const R csi = on ? Num<R>::sqrt(sm.Dg[i]) * dsq : (R)1;
      const R diag = csi + eps_dsq;
      R acc = on ? sm.x[i] - sm.mu[i] : (R)0;
      R yi = 0;
      for (int j = 0; j < d; ++j) {
        const R yj = __shfl_sync(0xffffffffu, acc / diag, j);
        const R csj = __shfl_sync(0xffffffffu, csi, j);
        if (i == j) yi = yj;
        if (on && i > j) acc = fma(-sm.Lt[cm_idx(i, j, d)] * csj, yj, acc);
      }
      const R nsq = warp_sum(on ? yi * yi : (R)0);
      const R r = (R)1 / (nsq + (R)1);
      const R zi = (R)2 * yi * r;
      const R zl = (nsq - (R)1) * r;
      const R vni = on ? xb[i] : (R)0;
      R vl = sm.scal[5];
      const R dot = warp_sum(on ? vni * zi : (R)0) + vl * zl;
      R vi = on ? fma(-dot, zi, vni) : (R)0;
      vl = fma(-dot, zl, vl);
      const R vsq = warp_sum(vi * vi) + vl * vl;
      const R rn = (R)1 / Num<R>::sqrt(vsq);
      vi *= rn;
      vl *= rn;
      if (on) { zc[i] = zi; vc[i] = vi; cs[i] = csi; }
      if (i == 0) {
        sm.scal[0] = zl;
        sm.scal[1] = vl;
        const R pe_z = U + (R)d * Num<R>::log((R)1 - zl);   // :228 (stored U == U(x(z)) up to round-off)
        sm.scal[2] = pe_z - Num<R>::log(sm.scal[6]);
      }
    }
    __syncthreads();
    // ---- shrinkage (:59-96); the control flow is uniform over the CTA
    const R t_pe = sm.scal[2];
    R theta = two_pi * sm.scal[7], th_min = theta - two_pi, th_max = theta;
    int iter = 0;
    R Un = 0;
    uint32_t cache[4];
    int cached_blk = -1;
    while (true) {
      R sn, cn;
      SinCos<R>::eval(theta, sn, cn);
      R om;
      const R pe = transformed(cn, sn, Un, om);
      const bool cont = (iter < kAsssMaxIter) && ((pe > t_pe) || (om < a.eps));
      if (!cont) break;
      if (theta < (R)0) th_min = theta; else th_max = theta;
      R un;
      if (EXTERNAL) {
        un = a.uniforms[(t * kAsssUniforms + 2 + iter) * C + c];
      } else {
        const int blk = 64 + (iter >> 2);
        if (blk != cached_blk) { rng.block((uint64_t)it, (uint32_t)blk, cache); cached_blk = blk; }
        un = (R)word_to_uniform(cache[iter & 3]);
      }
      theta = fma(un, th_max - th_min, th_min);
      ++iter;
    }
    if (iter >= kAsssMaxIter) {  // :94  give up: theta = 0 (stay at z)
      R om;
      transformed((R)1, (R)0, Un, om);
    }
    if (Num<R>::isnan(Un)) Un = Num<R>::inf();  // :244
    // ---- adaptation (:246-267)
    const int64_t n = (it < a.num_warmup) ? (it + 1) : (it + 1 - a.num_warmup);
    const R nf = (R)n;
    const bool n_is_one = (n == 1);
    const R gamma = n_is_one ? (R)1 : Num<R>::pow_neg(nf, a.lr_decay);
    const bool last = (t == a.n_steps - 1);
    __syncthreads();  // everyone is done with xb (x_base) before it becomes delta
    int ok_local = 1;
    for (int k = tid; k < d; k += NT) {
      const R xn = sm.xp[k];
      const R dl = xn - sm.mu[k];
      sm.x[k] = xn;
      if (ADAPT) sm.mu[k] = fma(gamma, dl, sm.mu[k]);
      sm.w[k] = dl;
      ok_local &= (Num<R>::abs(dl) < Num<R>::kBig) && (sm.Dg[k] > (R)0);
    }
    U = Un;
    const int ok = __syncthreads_and(ok_local) && !n_is_one && ADAPT;  // frozen (sample_Pnx): no update at all
    if (ADAPT && tid < 32) {
      R dn = 0;
      if (last) dn = warp_sum(tid < d ? sm.w[tid] * sm.w[tid] : (R)0);
      R ss = 0;
      if (ok) ss = last ? sweep_warp<R, true>(sm, d, gamma, (R)1, (R)1) : sweep_warp<R, false>(sm, d, gamma, (R)1, (R)1);
      if (last && tid == 0) sm.scal[3] = gamma * Num<R>::sqrt(dn) + Num<R>::sqrt(ss);  // :259-267
    }
    macc = fma((R)iter - macc, Num<R>::rcp(nf), macc);
    __syncthreads();
    if (last) asc = sm.scal[3];
    if (--until_collect == 0) {
      until_collect = a.thinning;
      if (a.out_z && lead)
        for (int k = tid; k < d; k += NT) a.out_z[(sidx * d + k) * C + c] = sm.x[k];
      if (a.out_pe && tid == 0 && lead) a.out_pe[sidx * C + c] = U;
      ++sidx;
    }
  }
  if (CL > 1) cooperative_groups::this_cluster().sync();  // nobody leaves while a peer may still write into its slots
  if (!lead) return;
  // ---- store
  __syncthreads();
  for (int k = tid; k < d; k += NT) st.z[k * C + c] = sm.x[k];
  if (tid == 0) st.pe[c] = U;
  if (!ADAPT) return;
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
  if (tid == 0) { st.macc[c] = macc; st.asc[c] = asc; }
}

#endif  // __CUDACC__
}  // namespace amcmc
