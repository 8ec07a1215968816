// diamonds_tc_adapt.cu -- diamonds on the tensor cores WITH per-chain adaptation: the full ARWMH.sample of the
// reference (python/kernels/arwmh.py:140-207) for tens of thousands of chains, the N-row likelihood on tcgen05
// (same GEMM formulation, operand tiles, TMEM epilogue and warp roles as diamonds_tc.cu) and every chain's own
// running mean, LDL^T proposal factor and step size updated between the tensor-core phases.
//
// Per-chain adaptation state cannot stay on chip here (128 chains x 1.4 KB per UMMA tile row block, four
// blocks per SM), so it lives in global memory in the struct-of-arrays layout of the ABI -- the chain index is
// fastest, so the 128 owner threads of a group touch it with fully coalesced accesses -- and is L2-resident
// (92 MB at 65,536 chains).  At every step boundary the owner thread of a chain makes ONE software-pipelined pass
// over its factor, column by column (the next column is loaded before the current one is stored):
//     rank-one update with delta = x_new - loc (Gill-Golub-Murray-Saunders recurrence, as arwmh_small.cuh)
//     + accumulation of the NEXT proposal  L~'(sqrt(D') .* z_next)  while the column is in registers,
// so the factor is read once and written once per step.  The helper warps run one step ahead and only
// produce the draws (z, u).  During the launch `scale` holds the factor in LDL^T form in place (slot (j,j) = D_j,
// slot (i,j) = L~_ij); two small kernels convert from / to the Cholesky form of the ABI around the launch.
#include <cmath>
#include <vector>
#include "diamonds_tc.cuh"

namespace amcmc {

struct TcAdaptParams {
  TcParams p;        // chains, tiles, positions, energies, draws, outputs (as the shared-state kernel)
  float* loc;        // [26][C]
  float* scale;      // [351][C]  LDL^T form during the launch
  float* lam;        // [C]
  float* asc;        // [C]
  int64_t num_warmup;
  float lr_decay, target, eps;
};

__device__ __forceinline__ int tri_f(int i, int j) { return i * (i + 1) / 2 + j; }

// in-place Cholesky <-> LDL^T of every chain's packed factor
__global__ void tc_chol_to_ldl_kernel(float* __restrict__ sc, int64_t C) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
#pragma unroll 1
  for (int j = 0; j < TC_D; ++j) {
    const float dg = sc[(int64_t)tri_f(j, j) * C + c];
    const float inv = 1.0f / dg;
    for (int i = j + 1; i < TC_D; ++i) sc[(int64_t)tri_f(i, j) * C + c] *= inv;
    sc[(int64_t)tri_f(j, j) * C + c] = dg * dg;
  }
}
__global__ void tc_ldl_to_chol_kernel(float* __restrict__ sc, int64_t C) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
#pragma unroll 1
  for (int j = 0; j < TC_D; ++j) {
    const float sd = sqrtf(sc[(int64_t)tri_f(j, j) * C + c]);
    for (int i = j + 1; i < TC_D; ++i) sc[(int64_t)tri_f(i, j) * C + c] *= sd;
    sc[(int64_t)tri_f(j, j) * C + c] = sd;
  }
}

// mean position over the chains -> reference point of the centred GEMM (float64 atomics, 26 values)
__global__ void tc_mean_kernel(const float* __restrict__ z, int64_t C, double* __restrict__ acc) {
  const int k = blockIdx.y;
  double s = 0;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < C; c += (int64_t)gridDim.x * blockDim.x)
    s += (double)z[(int64_t)k * C + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&acc[k], s);
}
__global__ void tc_mean_finish_kernel(const double* __restrict__ acc, int64_t C, float* __restrict__ loc, float* __restrict__ ident,
                                      float* __restrict__ zero) {
  const int t = threadIdx.x;
  if (t < TC_D) loc[t] = (float)(acc[t] / (double)C);
  for (int e = t; e < TC_NP; e += blockDim.x) ident[e] = 0.f;
  if (t == 0) zero[0] = 0.f;
}

// One pass over the chain's LDL^T factor in global memory (see the file header).
//   UPDATE : apply (1-gamma) L D L^T + gamma w w^T   (w is consumed);  otherwise the factor is only read
//   WANT   : accumulate |L' e^lam' - L e^lam|_F^2  (arwmh.py:197)
//   acc_i  = sum_{j<=i} L~'_ij sqrt(D'_j) zn_j     (zn: next draws in shared memory, stride TC_M; zero if !have_next)
template <bool UPDATE, bool WANT>
__device__ __forceinline__ float tc_column_pass(float* __restrict__ sc, int64_t C, int64_t c, const float* zn, bool have_next,
                                                float (&w)[TC_D], float gamma, float el_old, float el_new,
                                                float (&acc)[TC_D]) {
  float t = 1.f, ss = 0.f;
  const float omg = 1.f - gamma;
#pragma unroll
  for (int k = 0; k < TC_D; ++k) acc[k] = 0.f;
  float nxt[TC_D];  // prefetched column: nxt[j] = D_j, nxt[i > j] = L~_ij
#pragma unroll
  for (int i = 0; i < TC_D; ++i) nxt[i] = sc[(int64_t)tri_f(i, 0) * C + c];
#pragma unroll
  for (int j = 0; j < TC_D; ++j) {
    float cur[TC_D];
#pragma unroll
    for (int i = j; i < TC_D; ++i) cur[i] = nxt[i];
    if (j + 1 < TC_D) {
#pragma unroll
      for (int i = j + 1; i < TC_D; ++i) nxt[i] = sc[(int64_t)tri_f(i, j + 1) * C + c];  // in flight during column j
    }
    const float Dold = cur[j];
    float Dnew = Dold, coef = 0.f, wj = 0.f;
    if (UPDATE) {
      const bool pos = Dold > 0.f;  // a non-positive pivot leaves its column untouched
      const float Dj = omg * Dold;
      wj = w[j];
      const float cw = gamma * wj;
      const float g = fmaf(cw * wj, t, Dj);
      const float tr = __fdividef(t, g);
      if (pos) { coef = cw * tr; t = Dj * tr; Dnew = g; } else { wj = 0.f; }
      sc[(int64_t)tri_f(j, j) * C + c] = Dnew;
    }
    const float sn_raw = sqrtf(Dnew);
    const float yj = have_next ? sn_raw * zn[j * TC_M] : 0.f;
    acc[j] += yj;
    float so = 0.f, sn = 0.f;
    if (WANT) {
      so = sqrtf(Dold) * el_old;
      sn = sn_raw * el_new;
      const float dd = sn - so;
      ss = fmaf(dd, dd, ss);
    }
#pragma unroll
    for (int i = j + 1; i < TC_D; ++i) {
      const float Lo = cur[i];
      float Ln = Lo;
      if (UPDATE) {
        w[i] = fmaf(-wj, Lo, w[i]);
        Ln = fmaf(coef, w[i], Lo);
        sc[(int64_t)tri_f(i, j) * C + c] = Ln;
      }
      acc[i] = fmaf(Ln, yj, acc[i]);
      if (WANT) {
        const float df = fmaf(Ln, sn, -(Lo * so));
        ss = fmaf(df, df, ss);
      }
    }
  }
  return ss;
}

// proposal -> A' row (split bf16), scalar part of U', shadow position buffer.  Returns via references.
__device__ __forceinline__ void tc_emit_proposal(const float (&xp)[TC_D], const float* sRef, unsigned char* sA, int g, int row,
                                                 uint64_t* a_ready_g, double n_rows, double cst, double& Up_part, double& inv2var) {
  float dq = 0.f;
  uint16_t ak[TC_KP];
#pragma unroll
  for (int k = 0; k < TC_KC; ++k) {
    const float dlt = xp[k] - sRef[REF_Q + k];
    dq = fmaf(dlt, sRef[REF_G2 + k], dq);
    const uint16_t hi = f2bf(dlt);
    ak[k] = hi;
    ak[TC_KC + k] = f2bf(dlt - bf2f(hi));
    ak[2 * TC_KC + k] = hi;
  }
#pragma unroll
  for (int k = 3 * TC_KC; k < TC_KP; ++k) ak[k] = 0;
  unsigned char* arow = sA + g * TC_A_BYTES + (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
  for (int kc = 0; kc < TC_KP / 8; ++kc) {
    uint4 v;
    v.x = (uint32_t)ak[8 * kc] | ((uint32_t)ak[8 * kc + 1] << 16);
    v.y = (uint32_t)ak[8 * kc + 2] | ((uint32_t)ak[8 * kc + 3] << 16);
    v.z = (uint32_t)ak[8 * kc + 4] | ((uint32_t)ak[8 * kc + 5] << 16);
    v.w = (uint32_t)ak[8 * kc + 6] | ((uint32_t)ak[8 * kc + 7] << 16);
    *reinterpret_cast<uint4*>(arow + kc * (TC_M / 8) * 128) = v;
  }
  fence_proxy_async_smem();
  mbar_arrive(a_ready_g);
  float sb = 0.f;
#pragma unroll
  for (int k = 1; k < TC_KC; ++k) sb = fmaf(xp[k], xp[k], sb);
  const float s = xp[TC_D - 1];
  const float ti = (xp[0] - 8.f) * 0.1f, ts = __expf(s) * 0.1f;
  inv2var = 0.5 * exp(-2.0 * (double)s);
  Up_part = (double)(0.5f * sb + 2.f * log1pf(ti * ti * (1.f / 3.f)) + 2.f * log1pf(ts * ts * (1.f / 3.f))) +
            (n_rows - 1.0) * (double)s + cst + inv2var * (*reinterpret_cast<const double*>(sRef + REF_RSS) - (double)dq);
}

template <bool EXTERNAL>
__global__ void __launch_bounds__(TC_THREADS, 1) diamonds_tc_adapt_kernel(const TcAdaptParams ap) {
  const TcParams& p = ap.p;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sX = smem + TcSmem::OFF_X;
  unsigned char* sA = smem + TcSmem::OFF_A;
  float* sRef = reinterpret_cast<float*>(smem + TcSmem::OFF_REF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TcSmem::OFF_BAR);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + 2;
  uint64_t* acc_full = bars + 4;
  uint64_t* acc_empty = bars + 6;
  uint64_t* a_ready = bars + 8;
  uint64_t* v_full = bars + 8 + TC_GR;
  uint64_t* v_empty = bars + 8 + 2 * TC_GR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TcSmem::OFF_TMEM);
  float* sExch = reinterpret_cast<float*>(smem + TcSmem::OFF_EXCH);
  float* sV = reinterpret_cast<float*>(smem + TcSmem::OFF_V);  // [TC_GR][27][TC_M]: next draws z[26], u

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int per = p.n_groups / (int)gridDim.x, rem = p.n_groups % (int)gridDim.x;
  const int g_begin = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
  const int g_count = per + ((int)blockIdx.x < rem ? 1 : 0);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 32 * TC_EPI_WARPS);
    }
    for (int g = 0; g < TC_GR; ++g) {
      mbar_init(&a_ready[g], TC_M);
      mbar_init(&v_full[g], 32 * TC_HELP_WARPS);
      mbar_init(&v_empty[g], TC_M);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int e = tid; e < REF_FLOATS; e += TC_THREADS) sRef[e] = p.ref[e];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_rounds = (g_count + TC_GR - 1) / TC_GR;
  uint32_t x_it = 0, acc_it = 0, a_it = 0;

  for (int rnd = 0; rnd < n_rounds; ++rnd) {
    const int G = min(TC_GR, g_count - rnd * TC_GR);
    const int g0 = g_begin + rnd * TC_GR;

    if (warp == 0) {
      if (lane == 0) {  // ===== TMA producer =====
        for (int64_t st = 0; st < p.n_steps; ++st)
          for (int tile = 0; tile < p.n_tiles; ++tile, ++x_it) {
            const int s = x_it & 1;
            mbar_wait(&x_empty[s], ((x_it >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&x_full[s], TC_TILE_BYTES);
            tma_load_1d(sX + s * TC_TILE_BYTES, p.Xcanon + (size_t)tile * (TC_TILE_BYTES / 2), TC_TILE_BYTES, &x_full[s]);
          }
      }
    } else if (warp == 1) {
      if (lane == 0) {  // ===== MMA issuer =====
        const uint32_t idesc = make_idesc_bf16_f32(TC_M, TC_TILE_N);
        constexpr uint32_t a_kstride = (TC_M / 8) * 128, b_kstride = (TC_TILE_N / 8) * 128;
        for (int64_t st = 0; st < p.n_steps; ++st) {
          for (int tile = 0; tile < p.n_tiles; ++tile, ++x_it) {
            const int s = x_it & 1;
            mbar_wait(&x_full[s], (x_it >> 1) & 1);
            for (int g = 0; g < G; ++g, ++acc_it) {
              if (tile == 0) mbar_wait(&a_ready[g], (a_it + (uint32_t)st) & 1);
              const int b = acc_it & 1;
              mbar_wait(&acc_empty[b], ((acc_it >> 1) & 1) ^ 1);
              tc_fence_after();
              const uint32_t a_base = smem_u32(sA + g * TC_A_BYTES), b_base = smem_u32(sX + s * TC_TILE_BYTES);
#pragma unroll
              for (int ks = 0; ks < TC_KP / 16; ++ks) {
                const uint64_t da = make_smem_desc(a_base + 2 * ks * a_kstride, a_kstride, 128);
                const uint64_t db = make_smem_desc(b_base + 2 * ks * b_kstride, b_kstride, 128);
                umma_bf16(tmem_base + (uint32_t)(b * TC_TILE_N), da, db, idesc, ks > 0);
              }
              umma_commit(&acc_full[b]);
            }
            umma_commit(&x_empty[s]);
          }
        }
      }
    } else if (warp >= 2 + TC_EPI_WARPS) {
      // ===== helper warps: the draws of every chain, one step ahead (arwmh.py:162-165,174) =====
      const int ht = tid - 32 * (2 + TC_EPI_WARPS);
      for (int64_t st = 0; st < p.n_steps; ++st) {
        const int64_t it = p.i0 + st;
        for (int g = 0; g < G; ++g) {
          mbar_wait(&v_empty[g], ((a_it + (uint32_t)st) & 1) ^ 1);
          for (int row = ht; row < TC_M; row += 32 * TC_HELP_WARPS) {
            const int64_t c = (int64_t)(g0 + g) * TC_M + row;
            const int64_t cc = c < p.C ? c : (p.C - 1);
            float* vrow = sV + (size_t)g * 27 * TC_M + row;
            if (EXTERNAL) {
#pragma unroll
              for (int k = 0; k < TC_D; ++k) vrow[k * TC_M] = p.normals[(st * TC_D + k) * p.C + cc];
              vrow[26 * TC_M] = p.uniforms[st * p.C + cc];
            } else {
              float zz[TC_D], u;
              const Philox rng(p.seed, (uint64_t)(cc + p.chain_offset));
              philox_draws<float, TC_D>(rng, (uint64_t)it, zz, u);
#pragma unroll
              for (int k = 0; k < TC_D; ++k) vrow[k * TC_M] = zz[k];
              vrow[26 * TC_M] = u;
            }
          }
          mbar_arrive(&v_full[g]);
        }
      }
    } else {
      // ===== epilogue / sampler threads =====
      const int q4 = warp & 3;
      const int half = (warp - 2) >> 2;
      const int row = q4 * 32 + lane;
      const uint32_t t_lane = ((uint32_t)(q4 * 32) << 16) + (uint32_t)(half * 128);
      float Ucur[2] = {0.f, 0.f}, macc[2] = {0.f, 0.f}, lam[2] = {0.f, 0.f}, uacc[2] = {2.f, 2.f};
      double Up_part[2] = {0.0, 0.0}, inv2var[2] = {0.0, 0.0};
      int cur[2] = {0, 0};
      int64_t until_collect = p.collect_start + p.thinning;
      int64_t sidx = 0;

      // builds the proposal of step `st_next` for owned slot o from the current position, writes A', energy parts
      auto propose = [&](int o, int g, int64_t c, bool live, int64_t cc, const float (&acc)[TC_D], float el, uint32_t vphase) {
        const float* xsrc = cur[o] ? p.xprop : p.z;
        float* xdst = cur[o] ? p.z : p.xprop;
        const float* vrow = sV + (size_t)g * 27 * TC_M + row;
        float xp[TC_D];
#pragma unroll
        for (int i = 0; i < TC_D; ++i) {
          xp[i] = xsrc[(int64_t)i * p.C + cc] + fmaf(el, acc[i], ap.eps * vrow[i * TC_M]);  // arwmh.py:166-167
          if (live) xdst[(int64_t)i * p.C + c] = xp[i];
        }
        uacc[o] = vrow[26 * TC_M];
        mbar_arrive(&v_empty[g]);
        tc_emit_proposal(xp, sRef, sA, g, row, &a_ready[g], (double)p.n_rows, p.cst, Up_part[o], inv2var[o]);
        (void)vphase;
      };

      // ---- prologue: state of the owned chains, proposal of the first step (no update yet)
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const int g = 2 * half + o;
        if (g < G) {
          const int64_t c = (int64_t)(g0 + g) * TC_M + row;
          const bool live = c < p.C;
          const int64_t cc = live ? c : (p.C - 1);
          Ucur[o] = p.pe[cc]; macc[o] = p.macc[cc]; lam[o] = ap.lam[cc];
          mbar_wait(&v_full[g], a_it & 1);
          float w[TC_D], acc[TC_D];
#pragma unroll
          for (int k = 0; k < TC_D; ++k) w[k] = 0.f;
          tc_column_pass<false, false>(ap.scale, p.C, cc, sV + (size_t)g * 27 * TC_M + row, true, w, 0.f, 1.f, 1.f, acc);
          propose(o, g, c, live, cc, acc, __expf(lam[o]), 0);
        }
      }

      for (int64_t st = 0; st < p.n_steps; ++st) {
        const int64_t it = p.i0 + st;
        float rss[TC_GR];
#pragma unroll
        for (int g = 0; g < TC_GR; ++g) rss[g] = 0.f;
        // ---- likelihood: sum_n m_n^2 from the TMEM accumulators
        for (int tile = 0; tile < p.n_tiles; ++tile) {
#pragma unroll
          for (int g = 0; g < TC_GR; ++g) {
            if (g < G) {
              const int b = acc_it & 1;
              mbar_wait(&acc_full[b], (acc_it >> 1) & 1);
              tc_fence_after();
              const float ss = epilogue_sumsq_half(tmem_base + t_lane + (uint32_t)(b * TC_TILE_N));
              tc_fence_before();
              mbar_arrive(&acc_empty[b]);
              rss[g] += ss;
              ++acc_it;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < TC_GR; ++g)
          if (g < G && (g >> 1) != half) sExch[g * TC_M + row] = rss[g];
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        float mine[2];
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int g = 2 * half + o;
          mine[o] = (g < G) ? rss[g] + sExch[g * TC_M + row] : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");

        const bool collect_now = (--until_collect == 0);
        if (collect_now) until_collect = p.thinning;
        const bool last = (st == p.n_steps - 1);
        const int64_t n = (it < ap.num_warmup) ? (it + 1) : (it + 1 - ap.num_warmup);
        const float nf = (float)n;
        const float gamma = (n == 1) ? 1.f : Num<float>::pow_neg(nf, ap.lr_decay);
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int g = 2 * half + o;
          if (g < G) {
            const int64_t c = (int64_t)(g0 + g) * TC_M + row;
            const bool live = c < p.C;
            const int64_t cc = live ? c : (p.C - 1);
            // ---- accept / reject (arwmh.py:170-178)
            float Up = (float)(Up_part[o] + inv2var[o] * (double)mine[o]);
            if (Up != Up) Up = INFINITY;
            const float e = __expf(Ucur[o] - Up);
            const float alpha = (e > 1.f) ? 1.f : e;
            const bool accd = uacc[o] < alpha;
            macc[o] = fmaf(alpha - macc[o], Num<float>::rcp(nf), macc[o]);  // :185
            if (accd) { cur[o] ^= 1; Ucur[o] = Up; }
            const float* xs = cur[o] ? p.xprop : p.z;
            if (live) {
              if (p.out_acc) p.out_acc[st * p.C + c] = (uint8_t)accd;
              if (collect_now) {
                if (p.out_z) {
#pragma unroll
                  for (int k = 0; k < TC_D; ++k) p.out_z[(sidx * TC_D + k) * p.C + c] = xs[(int64_t)k * p.C + c];
                }
                if (p.out_pe) p.out_pe[sidx * p.C + c] = Ucur[o];
              }
            }
            // ---- adaptation (:188-193): mean, step size, rank-one factor update fused with the next proposal
            float w[TC_D], acc[TC_D];
            float dabs = 0.f;
#pragma unroll
            for (int k = 0; k < TC_D; ++k) {
              const float mu = ap.loc[(int64_t)k * p.C + cc];
              const float dl = xs[(int64_t)k * p.C + cc] - mu;
              if (live) ap.loc[(int64_t)k * p.C + c] = fmaf(gamma, dl, mu);
              w[k] = dl;
              dabs += fabsf(dl);
            }
            const bool ok = live && (n != 1) && (dabs < Num<float>::kBig);  // padding lanes must never write a factor
            const float el_old = __expf(lam[o]);
            const float lam_new = fmaf(gamma, alpha - ap.target, lam[o]);
            const float el_new = __expf(lam_new);
            lam[o] = lam_new;
            const bool have_next = !last;
            if (have_next) mbar_wait(&v_full[g], (a_it + (uint32_t)st + 1u) & 1);
            const float* zn = sV + (size_t)g * 27 * TC_M + row;
            float ss = 0.f;
            if (last) {
              ss = ok ? tc_column_pass<true, true>(ap.scale, p.C, cc, zn, false, w, gamma, el_old, el_new, acc)
                      : tc_column_pass<false, true>(ap.scale, p.C, cc, zn, false, w, gamma, el_old, el_new, acc);
              if (live) ap.asc[c] = sqrtf(ss);  // :197
            } else {
              if (ok) tc_column_pass<true, false>(ap.scale, p.C, cc, zn, true, w, gamma, el_old, el_new, acc);
              else tc_column_pass<false, false>(ap.scale, p.C, cc, zn, true, w, gamma, el_old, el_new, acc);
              propose(o, g, c, live, cc, acc, el_new, 0);
            }
          }
        }
        if (collect_now) ++sidx;
      }
      // ---- write back the per-chain scalars and the position if it ended in the shadow buffer
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const int g = 2 * half + o;
        if (g < G) {
          const int64_t c = (int64_t)(g0 + g) * TC_M + row;
          if (c < p.C) {
            p.pe[c] = Ucur[o];
            p.macc[c] = macc[o];
            ap.lam[c] = lam[o];
            if (cur[o]) {
#pragma unroll
              for (int k = 0; k < TC_D; ++k) p.z[(int64_t)k * p.C + c] = p.xprop[(int64_t)k * p.C + c];
            }
          }
        }
      }
    }
    a_it += (uint32_t)p.n_steps;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// from diamonds_tc.cu
int diamonds_tc_prepare(const amcmc_model* m, int64_t n_chains, TcParams* p, const amcmc_state* st, const amcmc_run_args* a);
void diamonds_tc_launch_ref(const amcmc_model* m, const float* loc, const float* scale, const float* lam, double eps, cudaStream_t s);

// Adaptive run on the tensor cores: the full ARWMH.sample for every chain of *st.
int run_diamonds_tc_adapt(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  DiamondsTcExtra* ex = (DiamondsTcExtra*)m->extra;
  if (!ex) { set_error("diamonds tensor-core path needs K = 25 predictors and fp32"); return AMCMC_ERR_UNSUPPORTED; }
  TcAdaptParams ap;
  int rc = diamonds_tc_prepare(m, st->n_chains, &ap.p, st, a);
  if (rc) return rc;
  ap.loc = (float*)st->loc;
  ap.scale = (float*)st->scale;
  ap.lam = (float*)st->log_step_size;
  ap.asc = (float*)st->as_change;
  ap.num_warmup = a->num_warmup;
  ap.lr_decay = (float)a->lr_decay;
  ap.target = (float)a->target_accept_prob;
  ap.eps = (float)a->eps;
  const int64_t C = st->n_chains;
  // reference point of the centred GEMM = current mean position of the batch
  if ((rc = check_cuda(cudaMemsetAsync(ex->mean_acc, 0, sizeof(double) * 32, s), "cudaMemsetAsync"))) return rc;
  tc_mean_kernel<<<dim3(64, TC_D), 256, 0, s>>>((const float*)st->z, C, ex->mean_acc);
  tc_mean_finish_kernel<<<1, 128, 0, s>>>(ex->mean_acc, C, ex->qmean, ex->ident, ex->zero);
  diamonds_tc_launch_ref(m, ex->qmean, ex->ident, ex->zero, 0.0, s);
  const unsigned gridc = (unsigned)((C + 127) / 128);
  tc_chol_to_ldl_kernel<<<gridc, 128, 0, s>>>(ap.scale, C);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = ap.p.n_groups < sms ? ap.p.n_groups : sms;
  if (a->rng_mode == AMCMC_RNG_EXTERNAL) {
    if ((rc = check_cuda(cudaFuncSetAttribute(diamonds_tc_adapt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::BYTES), "cudaFuncSetAttribute"))) return rc;
    diamonds_tc_adapt_kernel<true><<<grid, TC_THREADS, TcSmem::BYTES, s>>>(ap);
  } else {
    if ((rc = check_cuda(cudaFuncSetAttribute(diamonds_tc_adapt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::BYTES), "cudaFuncSetAttribute"))) return rc;
    diamonds_tc_adapt_kernel<false><<<grid, TC_THREADS, TcSmem::BYTES, s>>>(ap);
  }
  if ((rc = check_cuda(cudaGetLastError(), "diamonds_tc_adapt_kernel launch"))) return rc;
  tc_ldl_to_chol_kernel<<<gridc, 128, 0, s>>>(ap.scale, C);
  return check_cuda(cudaGetLastError(), "tc_ldl_to_chol_kernel launch");
}

}  // namespace amcmc
