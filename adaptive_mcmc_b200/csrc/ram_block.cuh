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

constexpr int kRamThreads = 512;  // two threads per row of the factor (row split at its middle 4-column chunk)

// ---- storage of the unit factor: ROW-major, 4-column chunks, rows padded so that 128-bit accesses of a warp
// (32 consecutive rows, each at its own right-to-left position) fall into distinct bank groups:
// the END chunk of consecutive rows advances by exactly one 4-element unit modulo 8.
AMCMC_HD int ram_nq(int i) { return (i + 3) >> 2; }  // 4-column chunks of row i (columns 0..i-1)
AMCMC_HD int ram_row_units(int i) {
  const int q = ram_nq(i), want = (1 - (ram_nq(i + 1) - q)) & 7;  // units(i) == want (mod 8), units(i) >= max(q, 1)
  const int base = q > 0 ? q : 1;
  return base + ((want - base) & 7);
}
AMCMC_HD int ram_total_units(int d) {
  int t = 0;
  for (int i = 0; i < d; ++i) t += ram_row_units(i);
  return t;
}

template <typename R> struct alignas(4 * sizeof(R) > 16 ? 16 : 4 * sizeof(R)) RVec4 { R a, b, c, e; };

// Shared-memory carve-up of the RAM kernel
template <typename R> struct RamSmem {
  R *beta, *pj, *pn;  // per-column parameters of the walk (chunk q = four consecutive columns = one vector load)
  R *x, *xp, *Dg, *v, *part0, *part1, *red, *scal, *Lt;
  int* rb;            // row base (in 4-element units) of every row
  __device__ RamSmem(unsigned char* base, int d) {
    const int dp = ((d + 3) & ~3) + 4;
    R* p = reinterpret_cast<R*>(base);
    beta = p; p += dp; pj = p; p += dp; pn = p; p += dp;
    x = p; p += dp; xp = p; p += dp; Dg = p; p += dp; v = p; p += dp; part0 = p; p += dp; part1 = p; p += dp;
    red = p; p += 32; scal = p; p += 8;
    rb = reinterpret_cast<int*>(p); p += (dp * sizeof(int) + sizeof(R) - 1) / sizeof(R);
    p = reinterpret_cast<R*>((reinterpret_cast<uintptr_t>(p) + 31) & ~(uintptr_t)31);
    Lt = p;
  }
  static size_t bytes(int d) {
    const int dp = ((d + 3) & ~3) + 4;
    return sizeof(R) * ((size_t)10 * dp + 40 + 8 + (size_t)4 * ram_total_units(d)) + 64;
  }
};

// One fused pass over the factor per MCMC step.  Thread (row i, segment) walks its half row RIGHT TO LEFT in
// 4-column chunks (one 128-bit load + one 128-bit store of the row, three 128-bit loads of the column parameters):
//     Lt_ij' = Lt_ij + beta_j w          (rank-one update; w = running suffix sum_{k>j}^{i} Lt_ik p_k)
//     w     += Lt_ij p_j
//     acc   += Lt_ij' p_j^next           (row sum of the NEXT proposal  v^next = Lt' p^next)
// so the factor is read once and written once per step (160 KB of SMEM traffic at d = 200, fp32).
template <typename R, bool MASK>
__device__ __forceinline__ void ram_chunk(RVec4<R>* Lrow, const RVec4<R>* b4, const RVec4<R>* p4, const RVec4<R>* n4,
                                          int q, int row, R& w, R& acc) {
  RVec4<R> L = Lrow[q];
  const RVec4<R> B = b4[q], P = p4[q], N = n4[q];
  const int j0 = 4 * q;
  if (!MASK || j0 + 3 < row) { const R n = fma(B.e, w, L.e); w = fma(L.e, P.e, w); acc = fma(n, N.e, acc); L.e = n; }
  if (!MASK || j0 + 2 < row) { const R n = fma(B.c, w, L.c); w = fma(L.c, P.c, w); acc = fma(n, N.c, acc); L.c = n; }
  if (!MASK || j0 + 1 < row) { const R n = fma(B.b, w, L.b); w = fma(L.b, P.b, w); acc = fma(n, N.b, acc); L.b = n; }
  { const R n = fma(B.a, w, L.a); w = fma(L.a, P.a, w); acc = fma(n, N.a, acc); L.a = n; }
  Lrow[q] = L;
}

template <class BM, typename R, bool EXTERNAL>
__global__ void __launch_bounds__(kRamThreads, 2) ram_block_kernel(const BM m, const StateView<R> st, const RunView<R> a,
                                                                   const int d) {
  constexpr int NT = kRamThreads;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RamSmem<R> sm(smem_raw, d);
  const int tid = threadIdx.x;
  const int seg = tid >> 8, row = tid & 255;
  const bool has_row = row < d;
  // this thread's chunks of row `row`: seg 1 = [qs, nq), seg 0 = [0, qs)
  const int nq = has_row ? ram_nq(row) : 0;
  const int qs = nq >> 1;
  const int q_hi = (seg == 1) ? nq : qs, q_lo = (seg == 1) ? qs : 0;
  int rbase = 0;
  for (int k = 0; k < row && k < d; ++k) rbase += ram_row_units(k);
  if (seg == 0 && has_row) sm.rb[row] = rbase;
  RVec4<R>* Lrow = reinterpret_cast<RVec4<R>*>(sm.Lt) + rbase;
  const RVec4<R>* b4 = reinterpret_cast<const RVec4<R>*>(sm.beta);
  const RVec4<R>* p4 = reinterpret_cast<const RVec4<R>*>(sm.pj);
  const RVec4<R>* n4 = reinterpret_cast<const RVec4<R>*>(sm.pn);

  const int64_t C = st.C, c = blockIdx.x;
  // ---- load: L (row-major packed with diagonal) -> unit factor in the padded row layout, Dg
  for (int k = tid; k < d; k += NT) {
    sm.x[k] = st.z[k * C + c];
    const R dg = st.scale[(int64_t)tri_full(k, k) * C + c];
    sm.Dg[k] = dg * dg;
    sm.v[k] = (R)1 / dg;
  }
  const int lt_floats = 4 * ram_total_units(d);
  for (int k = tid; k < lt_floats; k += NT) sm.Lt[k] = 0;  // padding must be finite
  __syncthreads();
  for (int e = tid; e < d * (d - 1) / 2; e += NT) {
    int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);
    while (i * (i - 1) / 2 > e) --i;
    while ((i + 1) * i / 2 <= e) ++i;
    const int j = e - i * (i - 1) / 2;
    sm.Lt[4 * sm.rb[i] + j] = st.scale[(int64_t)tri_full(i, j) * C + c] * sm.v[j];
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

  // the fused walk of this thread's half row; returns the row-sum contribution of the NEXT proposal
  auto walk = [&](R w) -> R {
    R acc = 0;
    int q = q_hi - 1;
    if (q >= q_lo && seg == 1) { ram_chunk<R, true>(Lrow, b4, p4, n4, q, row, w, acc); --q; }  // ragged right end
    for (; q >= q_lo; --q) ram_chunk<R, false>(Lrow, b4, p4, n4, q, row, w, acc);
    return acc;
  };

  // per-thread (tid < d) registers describing the CURRENT step's draw: z_j, z_j^2, prefix sum; |z|^2 in scal[2]
  R zj = 0, zsq = 0, pref = 0;
  // ---- prologue: draws of step 0, p^0 = sqrt(D) z^0, and one walk with beta = 0 to get v^0 = Lt p^0
  draw(0, a.i0, 0);
  __syncthreads();
  for (int k = tid; k < ((d + 3) & ~3) + 4; k += NT) { sm.beta[k] = 0; sm.pj[k] = 0; }
  if (tid < d) {
    zj = sm.v[tid];
    zsq = zj * zj;
    sm.pn[tid] = zj * Num<R>::sqrt(sm.Dg[tid]);  // p^0
    sm.part1[tid] = 0;
  } else if (tid < ((d + 3) & ~3) + 4) {
    sm.pn[tid] = 0;
  }
  pref = block_inclusive_scan<R, NT>(zsq, sm.red);
  if (tid == d - 1) sm.scal[2] = pref;
  __syncthreads();
  R* part_old = sm.part1;  // upper-segment sums belonging to the step being updated
  R* part_new = sm.part0;
  {
    const R acc = walk((R)0);
    if (has_row && seg == 1) part_new[row] = acc;
    __syncthreads();
    if (has_row && seg == 0) sm.v[row] = sm.pn[row] + acc + part_new[row];
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
      sm.beta[tid] = beta;
      sm.pj[tid] = pcur;
      sm.pn[tid] = zj * Num<R>::sqrt(Dnew);
      sm.Dg[tid] = Dnew;
    }
    pref = block_inclusive_scan<R, NT>(zsq, sm.red);
    if (tid == d - 1) sm.scal[2] = pref;
    __syncthreads();
    // ---- the fused pass: update the factor, accumulate the next proposal
    {
      const R w0 = has_row ? sm.pj[row] + (seg == 0 ? part_old[row] : (R)0) : (R)0;
      const R acc = walk(w0);
      if (has_row && seg == 1) part_new[row] = acc;
      __syncthreads();
      if (has_row && seg == 0) sm.v[row] = sm.pn[row] + acc + part_new[row];
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
    sm.v[k] = sd;
    st.scale[(int64_t)tri_full(k, k) * C + c] = sd;
  }
  __syncthreads();
  for (int e2 = tid; e2 < d * (d - 1) / 2; e2 += NT) {
    int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e2)) * 0.5f);
    while (i * (i - 1) / 2 > e2) --i;
    while ((i + 1) * i / 2 <= e2) ++i;
    const int j = e2 - i * (i - 1) / 2;
    st.scale[(int64_t)tri_full(i, j) * C + c] = sm.Lt[4 * sm.rb[i] + j] * sm.v[j];
  }
  if (tid == 0) { st.pe[c] = U; st.macc[c] = macc; }
}

#endif  // __CUDACC__
}  // namespace amcmc
