// ram_block.cuh -- robust adaptive Metropolis (Vihola 2012) for the correlated-Gaussian target of
// BASELINE.json configs[4] (d = 200): one CTA per chain, the 80 KB proposal factor lives in SHARED
// MEMORY across all fused steps (HBM streaming of the 164 KB state per step is the 3.98e7 steps/s
// roofline of SURVEY 8d; SMEM residency moves the bound to SMEM bandwidth / issue rate).
//
// NOT in the reference (spec: SURVEY 8a row 20, oracle: oracle/arwmh_numpy.py ram_step):
//   x' = x + L z;  alpha = min(1, pi(x')/pi(x));  L'L'^T = L (I + eta_n (alpha - alpha*) z z^T / |z|^2) L^T,
//   eta_n = min(1, d n^-lr_decay)
// With L = Lt diag(sqrt(D)) the rank-one vector is v = L z = Lt p, p = sqrt(D) .* z, so the forward
// substitution of the LDL^T update is trivial (Lt^-1 v = p) and the whole update is row-parallel:
//   b_{j+1} = 1 + c sum_{k<=j} z_k^2   (prefix sum; c = eta (alpha - alpha*) / |z|^2, 1 + c|z|^2 > 0 always)
//   D_j' = D_j b_{j+1} / b_j,   beta_j = c z_j / (sqrt(D_j) b_{j+1})
//   row i, right to left with the running suffix w = sum_{k>j}^{i} Lt_ik p_k:   Lt_ij' = Lt_ij + beta_j w
#pragma once
#include "arwmh_block.cuh"

namespace amcmc {

#ifdef __CUDACC__

// Gaussian potential with a banded precision factor: Pband[b][j] = P[j+b][j], b = 0..bw
template <typename R> struct GaussianBandModel {
  int d, bw;
  const R* __restrict__ Pband;  // [bw+1][d]
  template <int NT> __device__ R potential(const R* q, R* red) const {
    R acc = 0;
    for (int j = threadIdx.x; j < d; j += NT) {
      R v = 0;
      for (int b = 0; b <= bw; ++b)
        if (j + b < d) v = fma(q[j + b], __ldg(Pband + (size_t)b * d + j), v);
      acc = fma(v, v, acc);
    }
    return (R)0.5 * block_sum<R, NT>(acc, red);
  }
};

// inclusive scan over the CTA of one value per thread (NT = 256); `tmp` holds NT/32 values
template <typename R, int NT> __device__ __forceinline__ R block_inclusive_scan(R v, R* tmp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const R n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  __syncthreads();
  if (lane == 31) tmp[wid] = v;
  __syncthreads();
  R add = 0;
  for (int w = 0; w < wid; ++w) add += tmp[w];
  return v + add;
}

constexpr int kRamThreads = 512;  // two threads per row of the factor (row split at its midpoint)

// Shared-memory carve-up of the RAM kernel
template <typename R> struct RamSmem {
  R *x, *xp, *Dg, *v, *y, *part0, *part1, *red, *scal, *Lt;
  float4* pk;  // per column j: (beta_j, p_j of this step, p_j of the NEXT step, unused)  -- one 16-byte broadcast load
  __device__ RamSmem(unsigned char* base, int d) {
    const int dp = (d + 3) & ~3;
    pk = reinterpret_cast<float4*>(base);
    R* p = reinterpret_cast<R*>(base + sizeof(R) * 4 * dp);
    x = p; p += dp; xp = p; p += dp; Dg = p; p += dp; v = p; p += dp; y = p; p += dp;
    part0 = p; p += dp; part1 = p; p += dp; red = p; p += 32; scal = p; p += 8;
    Lt = p;
  }
  static size_t bytes(int d) {
    const int dp = (d + 3) & ~3;
    return sizeof(R) * ((size_t)11 * dp + 40 + (size_t)d * (d - 1) / 2 + 4);
  }
};

// One fused pass over the column-major packed unit factor per MCMC step.  Thread (row i, segment) walks its
// half row right to left; ALL lanes of a warp visit the same column j together (consecutive rows of a column
// are contiguous in the packed layout => conflict-free, and the per-column parameters are one broadcast
// 16-byte load).  For every element:
//     Lt_ij' = Lt_ij + beta_j w          (rank-one update; w = running suffix sum_{k>j}^{i} Lt_ik p_k)
//     w     += Lt_ij p_j
//     acc   += Lt_ij' p_j^next           (row sum of the NEXT proposal  v^next = Lt' p^next)
// so the factor is read once and written once per step (160 KB of SMEM traffic at d = 200, fp32).
template <typename R> struct RamPk { R beta, pj, pn, pad; };

template <typename R, bool PRED>
__device__ __forceinline__ void ram_walk(R* __restrict__ Lt, const R* pk4, int d, int row, int j_from, int j_to,
                                         int j_lo, int j_hi, R& w, R& acc) {
  // visits j = j_from-1, ..., j_to (descending); pk4 is the packed per-column parameter array (4 values each)
  const RamPk<R>* pk = reinterpret_cast<const RamPk<R>*>(pk4);
  int j = j_from - 1;
  int ad = colbase(j, d) + row - j - 1;  // address of (row, j); (row, j-1) sits d - 1 - j lower
  if (!PRED) {
    // four columns per trip: the four loads are issued before any store so that their latency overlaps
    for (; j - 3 >= j_to; j -= 4) {
      const int a0 = ad, a1 = a0 - (d - 1 - j), a2 = a1 - (d - j), a3 = a2 - (d + 1 - j);
      const R L0 = Lt[a0], L1 = Lt[a1], L2 = Lt[a2], L3 = Lt[a3];
      const RamPk<R> p0 = pk[j], p1 = pk[j - 1], p2 = pk[j - 2], p3 = pk[j - 3];
      const R n0 = fma(p0.beta, w, L0); w = fma(L0, p0.pj, w);
      const R n1 = fma(p1.beta, w, L1); w = fma(L1, p1.pj, w);
      const R n2 = fma(p2.beta, w, L2); w = fma(L2, p2.pj, w);
      const R n3 = fma(p3.beta, w, L3); w = fma(L3, p3.pj, w);
      Lt[a0] = n0; Lt[a1] = n1; Lt[a2] = n2; Lt[a3] = n3;
      acc = fma(n0, p0.pn, acc); acc = fma(n1, p1.pn, acc); acc = fma(n2, p2.pn, acc); acc = fma(n3, p3.pn, acc);
      ad = a3 - (d + 2 - j);
    }
  }
  for (; j >= j_to; --j) {
    if (!PRED || (j >= j_lo && j < j_hi)) {
      const R Lo = Lt[ad];
      const RamPk<R> p = pk[j];
      const R Ln = fma(p.beta, w, Lo);
      Lt[ad] = Ln;
      w = fma(Lo, p.pj, w);
      acc = fma(Ln, p.pn, acc);
    }
    ad -= d - 1 - j;
  }
}

template <class BM, typename R, bool EXTERNAL>
__global__ void __launch_bounds__(kRamThreads, 2) ram_block_kernel(const BM m, const StateView<R> st, const RunView<R> a,
                                                                   const int d) {
  constexpr int NT = kRamThreads;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RamSmem<R> sm(smem_raw, d);
  R* pk4 = reinterpret_cast<R*>(sm.pk);
  const int tid = threadIdx.x;
  const int seg = tid >> 8, row = tid & 255;
  const bool has_row = row < d;
  const int j_lo = has_row ? (seg == 0 ? 0 : row / 2) : 0;
  const int j_hi = has_row ? (seg == 0 ? row / 2 : row) : 0;  // [j_lo, j_hi)
  // warp-uniform loop bounds: union [u_lo, u_hi) and common part [c_lo, c_hi) of the lanes' segments
  int u_lo = (j_hi > j_lo) ? j_lo : 0x7fffffff, u_hi = (j_hi > j_lo) ? j_hi : 0;
  int c_lo = has_row ? j_lo : 0, c_hi = has_row ? j_hi : 0x7fffffff;  // rows beyond d do not constrain the common part
  if (!has_row) c_lo = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    u_lo = min(u_lo, __shfl_xor_sync(0xffffffffu, u_lo, o));
    u_hi = max(u_hi, __shfl_xor_sync(0xffffffffu, u_hi, o));
    c_lo = max(c_lo, __shfl_xor_sync(0xffffffffu, c_lo, o));
    c_hi = min(c_hi, __shfl_xor_sync(0xffffffffu, c_hi, o));
  }
  if (u_hi <= u_lo) { u_lo = 0; u_hi = 0; }
  // a warp containing out-of-range rows (row >= d) keeps everything predicated
  const bool all_rows = __all_sync(0xffffffffu, has_row);
  if (!all_rows || c_hi <= c_lo) { c_lo = u_lo; c_hi = u_lo; }  // empty common part
  c_hi = min(c_hi, u_hi);
  c_lo = max(c_lo, u_lo);

  const int64_t C = st.C, c = blockIdx.x;
  // ---- load: L (row-major packed with diagonal) -> Lt column-major packed, Dg
  for (int k = tid; k < d; k += NT) {
    sm.x[k] = st.z[k * C + c];
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
  R U = st.pe[c], macc = st.macc[c];
  __syncthreads();

  const Philox rng(a.seed, (uint64_t)(c + a.chain_offset));
  const int npair = (d + 1) / 2;
  int64_t until_collect = a.collect_start + a.thinning;
  int64_t sidx = 0;

  // draws of iteration `it` -> sm.v (temporarily z) and sm.scal[slot]
  auto draw = [&](int64_t t, int64_t it, int slot) {
    if (EXTERNAL) {
      for (int k = tid; k < d; k += NT) sm.v[k] = a.normals[(t * d + k) * C + c];
      if (tid == 0) sm.scal[slot] = a.uniforms[t * C + c];
    } else {
      for (int p = tid; p <= npair; p += NT) {
        uint32_t o[4];
        if (p < npair) {
          rng.block((uint64_t)it, (uint32_t)(p >> 1), o);
          float z0, z1;
          box_muller(o[(p & 1) * 2], o[(p & 1) * 2 + 1], z0, z1);
          sm.v[2 * p] = (R)z0;
          if (2 * p + 1 < d) sm.v[2 * p + 1] = (R)z1;
        } else {
          rng.block((uint64_t)it, (uint32_t)((2 * npair) >> 2), o);
          sm.scal[slot] = (R)word_to_uniform(o[(2 * npair) & 3]);
        }
      }
    }
  };

  // per-thread (tid < d) registers describing the CURRENT step's draw: z_j, z_j^2, prefix sum, |z|^2 in scal[2]
  R zj = 0, zsq = 0, pref = 0;
  // ---- prologue: draws of step 0, p^0 = sqrt(D) z^0, and one walk with beta = 0 to get v^0 = Lt p^0
  draw(0, a.i0, 0);
  __syncthreads();
  if (tid < d) {
    zj = sm.v[tid];
    zsq = zj * zj;
    pk4[4 * tid] = 0;                                      // beta
    pk4[4 * tid + 1] = 0;                                  // p of a (non-existent) previous step
    pk4[4 * tid + 2] = zj * Num<R>::sqrt(sm.Dg[tid]);      // p^0
    sm.part1[tid] = 0;
  }
  pref = block_inclusive_scan<R, NT>(zsq, sm.red);
  if (tid == d - 1) sm.scal[2] = pref;
  __syncthreads();
  R* part_old = sm.part1;  // upper-segment sums belonging to the step being updated
  R* part_new = sm.part0;
  {
    R w = 0, acc = 0;
    ram_walk<R, true>(sm.Lt, pk4, d, row, u_hi, c_hi, j_lo, j_hi, w, acc);
    ram_walk<R, false>(sm.Lt, pk4, d, row, c_hi, c_lo, j_lo, j_hi, w, acc);
    ram_walk<R, true>(sm.Lt, pk4, d, row, c_lo, u_lo, j_lo, j_hi, w, acc);
    if (has_row && seg == 1) part_new[row] = acc;
    __syncthreads();
    if (has_row && seg == 0) sm.v[row] = pk4[4 * row + 2] + acc + part_new[row];
    __syncthreads();
  }
  { R* tmp = part_old; part_old = part_new; part_new = tmp; }

  for (int64_t t = 0; t < a.n_steps; ++t) {
    const int64_t it = a.i0 + t;
    const int uslot = (int)(t & 1);  // uniform of step t lives in scal[uslot]
    // ---- proposal, potential, accept
    if (tid < d) sm.xp[tid] = sm.x[tid] + sm.v[tid];
    __syncthreads();
    R Up = m.template potential<NT>(sm.xp, sm.red);
    if (Num<R>::isnan(Up)) Up = Num<R>::inf();
    const R e = Num<R>::exp(U - Up);
    const R alpha = (e > (R)1) ? (R)1 : e;
    const bool accd = sm.scal[uslot] < alpha;
    const R zz = sm.scal[2];
    if (accd) {
      if (tid < d) sm.x[tid] = sm.xp[tid];
      U = Up;
    }
    if (a.out_acc && tid == 0) a.out_acc[t * C + c] = (uint8_t)accd;
    const int64_t n = (it < a.num_warmup) ? (it + 1) : (it + 1 - a.num_warmup);
    const R nf = (R)n;
    macc = fma(alpha - macc, Num<R>::rcp(nf), macc);
    // ---- coefficients of the rank-one update.  Vihola (2012): eta_n = min(1, d n^-lr_decay)
    const R eta_raw = (n == 1) ? (R)d : (R)d * Num<R>::pow_neg(nf, a.lr_decay);
    const R eta = eta_raw < (R)1 ? eta_raw : (R)1;
    R cfac = eta * (alpha - a.target) / zz;
    const bool ok = (zz > (R)0) && (cfac == cfac) && (Num<R>::abs(cfac) < Num<R>::kBig);
    if (!ok) cfac = 0;  // beta = 0, D' = D: the walk below still produces the next proposal
    R Dnew = 0, beta = 0, pcur = 0;
    if (tid < d) {
      const R b_next = fma(cfac, pref, (R)1);       // b_{j+1}
      const R b_cur = fma(cfac, pref - zsq, (R)1);  // b_j
      const R Dj = sm.Dg[tid];
      const R sq = Num<R>::sqrt(Dj);
      Dnew = Dj * b_next / b_cur;
      beta = cfac * zj / (sq * b_next);
      pcur = zj * sq;
    }
    __syncthreads();  // everyone has read scal[2] / v / xp of this step
    // ---- draws of the NEXT step (state independent), p^next = sqrt(D') z^next
    if (t + 1 < a.n_steps) draw(t + 1, it + 1, uslot ^ 1);  // (uniform branch; the last step needs no next proposal)
    __syncthreads();
    if (tid < d) {
      zj = sm.v[tid];
      zsq = zj * zj;
      pk4[4 * tid] = beta;
      pk4[4 * tid + 1] = pcur;
      pk4[4 * tid + 2] = zj * Num<R>::sqrt(Dnew);
      sm.Dg[tid] = Dnew;
    }
    pref = block_inclusive_scan<R, NT>(zsq, sm.red);
    if (tid == d - 1) sm.scal[2] = pref;
    __syncthreads();
    // ---- the fused pass: update the factor, accumulate the next proposal
    {
      R w = has_row ? pk4[4 * row + 1] + (seg == 0 ? part_old[row] : (R)0) : (R)0;
      R acc = 0;
      ram_walk<R, true>(sm.Lt, pk4, d, row, u_hi, c_hi, j_lo, j_hi, w, acc);
      ram_walk<R, false>(sm.Lt, pk4, d, row, c_hi, c_lo, j_lo, j_hi, w, acc);
      ram_walk<R, true>(sm.Lt, pk4, d, row, c_lo, u_lo, j_lo, j_hi, w, acc);
      if (has_row && seg == 1) part_new[row] = acc;
      __syncthreads();
      if (has_row && seg == 0) sm.v[row] = pk4[4 * row + 2] + acc + part_new[row];
    }
    { R* tmp = part_old; part_old = part_new; part_new = tmp; }
    __syncthreads();
    if (--until_collect == 0) {
      until_collect = a.thinning;
      if (a.out_z)
        for (int k = tid; k < d; k += NT) a.out_z[(sidx * d + k) * C + c] = sm.x[k];
      if (a.out_pe && tid == 0) a.out_pe[sidx * C + c] = U;
      ++sidx;
    }
  }
  // ---- store
  __syncthreads();
  for (int k = tid; k < d; k += NT) {
    st.z[k * C + c] = sm.x[k];
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
  if (tid == 0) { st.pe[c] = U; st.macc[c] = macc; }
}

#endif  // __CUDACC__
}  // namespace amcmc
