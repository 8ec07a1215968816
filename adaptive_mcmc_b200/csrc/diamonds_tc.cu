// diamonds_tc.cu -- diamonds (N = 5000-row linear regression, d = 26) many-chain Metropolis on the
// 5th-gen tensor cores: the X*beta likelihood of 128-chain groups is a dense contraction executed by
// tcgen05.mma (UTCHMMA) with TMA-staged (cp.async.bulk, UBLKCP) design-matrix tiles and fp32
// accumulators in TMEM; the epilogue reads TMEM (tcgen05.ld) and reduces sum-of-squares per chain.
//
// Model: python/scripts/run_diamonds_lr_decay.py:24-40.  Sampler step: python/kernels/arwmh.py:161-178
// with the adaptation state SHARED by all chains of the launch (the frozen kernel of sample_Pnx,
// arwmh.py:230-249, and the pooled-adaptation mode of BASELINE.json configs[3]).
//
// Numerics.  With a shared reference point q_ref (the pooled mean), Delta = (I', b') - (I_ref, b_ref):
//     RSS(q') = RSS_ref - 2 Delta.g + sum_n m_n^2,     m = [1 | Xc] Delta,   g = [1|Xc]^T r_ref
// RSS_ref and g come from the fp64 Gram matrix (tiny, recomputed per window); only sum_n m_n^2 needs
// the N-row contraction and it is a sum of squares of GEMM outputs (no cancellation).  Operands are
// split bf16 hi/lo and the three significant cross terms are concatenated along K:
//     A' = [D_hi | D_lo | D_hi | 0]   (128 chains x 80)      B' = [X_hi | X_hi | X_lo | 0]   (rows x 80)
// so one K = 80 GEMM (5 UMMA k-steps of 16) yields ~16-bit-mantissa products with fp32 accumulation.
//
// CTA = 384 threads: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-9 = 256 epilogue /
// sampler threads, warps 10-11 = helpers that compute the state-independent half of the NEXT step's
// proposal (draws, v = S z) while the tensor pipe is busy, so that the step boundary only costs
// x' = x + v, the bf16 split and ten 16-byte stores per chain.  Thread (quarter q, lane, half h) <-> TMEM lane (= chain row) 32q+lane; it drains
// columns [128h, 128h+128) of every accumulator tile (4 tcgen05.ld in flight before one wait) and owns
#include <cmath>
#include <vector>
#include "diamonds_tc.cuh"

namespace amcmc {

// per-window reference block: q_ref, 2g, RSS_ref (fp64 Gram algebra), S = e^lam L + eps I
template <typename R>
__global__ void diamonds_tc_ref_kernel(const double* __restrict__ gram, const R* __restrict__ loc,
                                       const R* __restrict__ scale_packed, const R* __restrict__ log_step, double eps,
                                       float* __restrict__ ref) {
  __shared__ double q[TC_KC];
  __shared__ double Gq[TC_KC];
  const int t = threadIdx.x;
  const double* G = gram;
  const double* h = gram + TC_KC * TC_KC;
  const double yy = gram[TC_KC * TC_KC + TC_KC];
  if (t < TC_KC) q[t] = (double)loc[t];
  __syncthreads();
  if (t < TC_KC) {
    double s = 0;
    for (int k = 0; k < TC_KC; ++k) s += G[t * TC_KC + k] * q[k];
    Gq[t] = s;
    ref[REF_G2 + t] = (float)(2.0 * (h[t] - s));
  }
  if (t < TC_D) ref[REF_Q + t] = (float)loc[t];
  __syncthreads();
  if (t == 0) {
    double rss = yy;
    for (int k = 0; k < TC_KC; ++k) rss += q[k] * (Gq[k] - 2.0 * h[k]);
    *reinterpret_cast<double*>(ref + REF_RSS) = rss;
  }
  const double el = exp((double)log_step[0]);
  for (int e = t; e < TC_NP; e += blockDim.x) {
    int i = (int)((sqrtf(1.0f + 8.0f * (float)e) - 1.0f) * 0.5f);
    while (i * (i + 1) / 2 > e) --i;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    const int j = e - i * (i + 1) / 2;
    ref[REF_S + i * REF_SLD + j] = (float)((double)scale_packed[e] * el + (i == j ? eps : 0.0));
  }
  for (int e = t; e < TC_D * REF_SLD; e += blockDim.x)
    if (e % REF_SLD > e / REF_SLD) ref[REF_S + e] = 0.f;
}

template <bool EXTERNAL>
__global__ void __launch_bounds__(TC_THREADS, 1) diamonds_tc_kernel(const TcParams p) {
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
  float* sV = reinterpret_cast<float*>(smem + TcSmem::OFF_V);  // [TC_GR][27][TC_M]: proposal increments + uniform

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // groups owned by this CTA: contiguous range, balanced to within one group (host guarantees grid <= n_groups)
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
  uint32_t x_it = 0, acc_it = 0, a_it = 0;  // pipeline iteration counters persist across rounds

  for (int rnd = 0; rnd < n_rounds; ++rnd) {
    const int G = min(TC_GR, g_count - rnd * TC_GR);
    const int g0 = g_begin + rnd * TC_GR;

    if (warp == 0) {
      // ===== TMA producer: stream the X' tiles, one pass per MCMC step =====
      if (lane == 0) {
        for (int64_t st = 0; st < p.n_steps; ++st)
          for (int tile = 0; tile < p.n_tiles; ++tile, ++x_it) {
            const int s = x_it & 1;
            mbar_wait(&x_empty[s], ((x_it >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&x_full[s], TC_TILE_BYTES);
            tma_load_1d(sX + s * TC_TILE_BYTES, p.Xcanon + (size_t)tile * (TC_TILE_BYTES / 2), TC_TILE_BYTES, &x_full[s]);
          }
      }
    } else if (warp == 1) {
      // ===== MMA issuer (one thread) =====
      if (lane == 0) {
        const uint32_t idesc = make_idesc_bf16_f32(TC_M, TC_TILE_N);
        constexpr uint32_t a_kstride = (TC_M / 8) * 128, b_kstride = (TC_TILE_N / 8) * 128;
        const uint64_t da0 = make_smem_desc(smem_u32(sA), a_kstride, 128), db0 = make_smem_desc(smem_u32(sX), b_kstride, 128);
        const uint32_t da_lo0 = (uint32_t)da0, da_hi = (uint32_t)(da0 >> 32), db_lo0 = (uint32_t)db0, db_hi = (uint32_t)(db0 >> 32);
        constexpr uint32_t kAChunk = (2 * a_kstride) >> 4, kBChunk = (2 * b_kstride) >> 4;  // one K = 16 chunk, in descriptor units
        for (int64_t st = 0; st < p.n_steps; ++st) {
          for (int tile = 0; tile < p.n_tiles; ++tile, ++x_it) {
            const int s = x_it & 1;
            mbar_wait(&x_full[s], (x_it >> 1) & 1);
            for (int g = 0; g < G; ++g, ++acc_it) {
              if (tile == 0) mbar_wait(&a_ready[g], (a_it + (uint32_t)st) & 1);
              const int b = acc_it & 1;
              mbar_wait(&acc_empty[b], ((acc_it >> 1) & 1) ^ 1);
              tc_fence_after();
              // hoisted base descriptors + chunk offsets in the low word: tcgen05.mma issue blocks the issuing thread, so its
              // own instruction stream is tensor-pipe idle time (scripts/probes/umma_bench.cu: 128 vs 232 cycles per MMA)
              const uint32_t a_lo = da_lo0 + (uint32_t)g * (TC_A_BYTES >> 4), b_lo = db_lo0 + (uint32_t)s * (TC_TILE_BYTES >> 4);
              const uint32_t d = tmem_base + (uint32_t)(b * TC_TILE_N);
              umma_bf16_lean<false>(d, a_lo, da_hi, b_lo, db_hi, idesc);
#pragma unroll
              for (int ks = 1; ks < TC_KP / 16; ++ks)
                umma_bf16_lean<true>(d, a_lo + ks * kAChunk, da_hi, b_lo + ks * kBChunk, db_hi, idesc);
              umma_commit(&acc_full[b]);
            }
            umma_commit(&x_empty[s]);
          }
        }
      }
    } else if (warp >= 2 + TC_EPI_WARPS) {
      // ===== helper warps: state-independent half of the proposal, one step ahead of the chains =====
      // v = (e^lam L + eps I) z  and  u   (arwmh.py:162-166,174) for every chain of the resident groups,
      // computed while the tensor pipe works on the previous step; handed over through shared memory.
      const int ht = tid - 32 * (2 + TC_EPI_WARPS);  // 0 .. 32*TC_HELP_WARPS-1
      const float* S = sRef + REF_S;
      for (int64_t st = 0; st < p.n_steps; ++st) {
        const int64_t it = p.i0 + st;
        for (int g = 0; g < G; ++g) {
          mbar_wait(&v_empty[g], ((a_it + (uint32_t)st) & 1) ^ 1);
          for (int row = ht; row < TC_M; row += 32 * TC_HELP_WARPS) {
            const int64_t c = (int64_t)(g0 + g) * TC_M + row;
            const int64_t cc = c < p.C ? c : (p.C - 1);
            float z[REF_SLD], u;
            if (EXTERNAL) {
#pragma unroll
              for (int k = 0; k < TC_D; ++k) z[k] = p.normals[(st * TC_D + k) * p.C + cc];
              u = p.uniforms[st * p.C + cc];
            } else {
              float zz[TC_D];
              const Philox rng(p.seed, (uint64_t)(cc + p.chain_offset));
              philox_draws<float, TC_D>(rng, (uint64_t)it, zz, u);
#pragma unroll
              for (int k = 0; k < TC_D; ++k) z[k] = zz[k];
            }
#pragma unroll
            for (int k = TC_D; k < REF_SLD; ++k) z[k] = 0.f;
            float* vrow = sV + (size_t)g * 27 * TC_M + row;
#pragma unroll
            for (int i = 0; i < TC_D; ++i) {
              float acc = 0.f;
#pragma unroll
              for (int j4 = 0; j4 <= i / 4; ++j4) {  // 16-byte broadcast loads of the shared proposal factor
                const float4 s4 = *reinterpret_cast<const float4*>(S + i * REF_SLD + 4 * j4);
                acc = fmaf(s4.x, z[4 * j4], acc);
                acc = fmaf(s4.y, z[4 * j4 + 1], acc);
                acc = fmaf(s4.z, z[4 * j4 + 2], acc);
                acc = fmaf(s4.w, z[4 * j4 + 3], acc);
              }
              vrow[i * TC_M] = acc;
            }
            vrow[26 * TC_M] = u;
          }
          mbar_arrive(&v_full[g]);
        }
      }
    } else {
      // ===== epilogue / sampler threads =====
      const int q4 = warp & 3;             // TMEM lane quarter this warp may access
      const int half = (warp - 2) >> 2;    // which 128 accumulator columns this thread drains
      const int row = q4 * 32 + lane;      // chain row within a group
      const uint32_t t_lane = ((uint32_t)(q4 * 32) << 16) + (uint32_t)(half * 128);
      // per owned group (g = 2*half + o): energy, position-buffer selector, acceptance sum live in registers
      float Ucur[2], macc_sum[2];
      int cur[2];
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const int g = 2 * half + o;
        const int64_t c = (int64_t)(g0 + g) * TC_M + row;
        Ucur[o] = (g < G && c < p.C) ? p.pe[c] : 0.f;
        macc_sum[o] = 0.f;
        cur[o] = 0;
      }
      int64_t until_collect = p.collect_start + p.thinning;
      int64_t sidx = 0;

      for (int64_t st = 0; st < p.n_steps; ++st) {
        // The handful of scalar terms of U' is assembled in float64: N*s and the normalisation constant are
        // O(1e4) and would cost ~1e-3 absolute in fp32 (40 DFMA-class ops per chain-step: negligible).
        double Up_part[2], inv2var[2];  // U' = Up_part + inv2var * max(RSS_ref - dq + sum m^2, 0)
        float dqv[2];
        float uacc[2], rss[TC_GR];
#pragma unroll
        for (int g = 0; g < TC_GR; ++g) rss[g] = 0.f;
        // ---- proposals (arwmh.py:167): x' = x + v; A' rows; scalar part of U'
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int g = 2 * half + o;
          Up_part[o] = 0.0; inv2var[o] = 0.0; dqv[o] = 0.f; uacc[o] = 2.f;
          if (g < G) {
            const int64_t c = (int64_t)(g0 + g) * TC_M + row;
            const bool live = c < p.C;
            const int64_t cc = live ? c : (p.C - 1);
            const float* xsrc = cur[o] ? p.xprop : p.z;
            float* xdst = cur[o] ? p.z : p.xprop;
            float xp[TC_D];
#pragma unroll
            for (int i = 0; i < TC_D; ++i) xp[i] = xsrc[(int64_t)i * p.C + cc];
            mbar_wait(&v_full[g], (a_it + (uint32_t)st) & 1);
            const float* vrow = sV + (size_t)g * 27 * TC_M + row;
#pragma unroll
            for (int i = 0; i < TC_D; ++i) {
              xp[i] += vrow[i * TC_M];
              if (live) xdst[(int64_t)i * p.C + c] = xp[i];  // the proposal lives in the other position buffer
            }
            uacc[o] = vrow[26 * TC_M];
            mbar_arrive(&v_empty[g]);
            float dq = 0.f;  // Delta . 2g
            uint16_t ak[TC_KP];  // this chain's row of A' = [D_hi | D_lo | D_hi | 0]
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
            // one 16-byte store per 8-element K chunk (= one core-matrix row of the canonical layout)
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
            mbar_arrive(&a_ready[g]);
            float sb = 0.f;
#pragma unroll
            for (int k = 1; k < TC_KC; ++k) sb = fmaf(xp[k], xp[k], sb);
            const float s = xp[TC_D - 1];
            const float ti = (xp[0] - 8.f) * 0.1f, ts = __expf(s) * 0.1f;
            // e^{-2s}/2 multiplies RSS ~ 75 into a ~2.5e3 term: an SFU exp (2^-21) would cost ~1e-3 absolute
            inv2var[o] = 0.5 * exp(-2.0 * (double)s);
            // U' = 1/2 sum b^2 + 2 log1p(ti^2/3) + 2 log1p(ts^2/3) + (N-1) s + cst + e^{-2s}/2 (RSS_ref - 2 D.g + sum m^2)
            Up_part[o] = (double)(0.5f * sb + 2.f * log1pf(ti * ti * (1.f / 3.f)) + 2.f * log1pf(ts * ts * (1.f / 3.f))) +
                         ((double)p.n_rows - 1.0) * (double)s + p.cst;
            dqv[o] = dq;
          }
        }
        // ---- likelihood: sum_n m_n^2 from the TMEM accumulators (this thread's 128 columns of every tile)
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
        // ---- exchange the column halves: hand the partial sums of the groups I do not own to my partner
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
        // ---- accept / reject (arwmh.py:170-178): flip the position-buffer selector instead of copying
        const bool collect_now = (--until_collect == 0);
        if (collect_now) until_collect = p.thinning;
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int g = 2 * half + o;
          if (g < G) {
            const int64_t c = (int64_t)(g0 + g) * TC_M + row;
            if (c < p.C) {
              // RSS is a sum of squares: assembled first and clamped at 0, so that the cancellation error of a
              // far-away proposal (|D| in the hundreds, e^{-2s} ~ 1e300) cannot turn into a large NEGATIVE energy
              const double rss = (*reinterpret_cast<const double*>(sRef + REF_RSS) - (double)dqv[o]) + (double)mine[o];
              float Up = (float)(Up_part[o] + inv2var[o] * (rss > 0.0 ? rss : 0.0));
              if (Up != Up) Up = INFINITY;
              const float e = __expf(Ucur[o] - Up);
              const float alpha = (e > 1.f) ? 1.f : e;
              const bool acc = uacc[o] < alpha;
              macc_sum[o] += alpha;
              if (acc) { cur[o] ^= 1; Ucur[o] = Up; }
              if (p.out_acc) p.out_acc[st * p.C + c] = (uint8_t)acc;
              if (collect_now) {
                if (p.out_z) {
                  const float* xs = cur[o] ? p.xprop : p.z;
#pragma unroll
                  for (int k = 0; k < TC_D; ++k) p.out_z[(sidx * TC_D + k) * p.C + c] = xs[(int64_t)k * p.C + c];
                }
                if (p.out_pe) p.out_pe[sidx * p.C + c] = Ucur[o];
              }
            }
          }
        }
        if (collect_now) ++sidx;
      }
      // ---- write back: energy, mean acceptance, and the position if it ended in the shadow buffer
      const float inv_n = 1.f / (float)p.n_steps;
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const int g = 2 * half + o;
        if (g < G) {
          const int64_t c = (int64_t)(g0 + g) * TC_M + row;
          if (c < p.C) {
            p.pe[c] = Ucur[o];
            p.macc[c] = macc_sum[o] * inv_n;
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

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static uint16_t host_f2bf(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  const uint32_t r = 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)((u + r) >> 16);
}
static float host_bf2f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float x;
  memcpy(&x, &u, 4);
  return x;
}

// Build the tensor-core operands of the diamonds design matrix (called from create_diamonds for K = 25).
int create_diamonds_tc(amcmc_model* m, const double* X, int64_t n, int K, const double* Y) {
  if (K != TC_KC || m->dtype != AMCMC_F32) return AMCMC_OK;  // tensor path compiled for the posteriordb shape, fp32
  const int kc = K - 1;
  std::vector<double> mean(kc, 0.0);
  for (int k = 0; k < kc; ++k) {
    for (int64_t r = 0; r < n; ++r) mean[k] += X[r * K + 1 + k];
    mean[k] /= (double)n;
  }
  const int n_tiles = (int)((n + TC_TILE_N - 1) / TC_TILE_N);
  std::vector<uint16_t> canon((size_t)n_tiles * (TC_TILE_BYTES / 2), 0);
  std::vector<uint16_t> canon64((size_t)n_tiles * (TC_TILE_N * 64), 0);
  std::vector<double> gram(TC_KC * TC_KC + TC_KC + 1, 0.0);
  std::vector<double> x1(TC_KC);
  for (int64_t r = 0; r < n; ++r) {
    x1[0] = 1.0;
    for (int k = 0; k < kc; ++k) x1[1 + k] = X[r * K + 1 + k] - mean[k];
    for (int a = 0; a < TC_KC; ++a) {
      for (int b = 0; b < TC_KC; ++b) gram[a * TC_KC + b] += x1[a] * x1[b];
      gram[TC_KC * TC_KC + a] += x1[a] * Y[r];
    }
    gram[TC_KC * TC_KC + TC_KC] += Y[r] * Y[r];
    const int tile = (int)(r / TC_TILE_N), rr = (int)(r % TC_TILE_N);
    uint16_t* t = canon.data() + (size_t)tile * (TC_TILE_BYTES / 2);
    uint16_t* t64 = canon64.data() + (size_t)tile * (TC_TILE_N * 64);
    for (int k = 0; k < TC_KC; ++k) {
      const float xf = (float)x1[k];
      const uint16_t hi = host_f2bf(xf);
      const uint16_t lo = host_f2bf((float)(x1[k] - (double)host_bf2f(hi)));
      t[canon_off(rr, k, TC_TILE_N) / 2] = hi;               // pairs with D_hi
      t[canon_off(rr, TC_KC + k, TC_TILE_N) / 2] = hi;       // pairs with D_lo
      t[canon_off(rr, 2 * TC_KC + k, TC_TILE_N) / 2] = lo;   // pairs with D_hi
      t64[canon_off(rr, k, TC_TILE_N) / 2] = hi;
      t64[canon_off(rr, 32 + k, TC_TILE_N) / 2] = lo;
    }
  }
  DiamondsTcExtra* ex = (DiamondsTcExtra*)calloc(1, sizeof(DiamondsTcExtra));
  ex->n_tiles = n_tiles;
  int rc;
  if ((rc = check_cuda(cudaMalloc(&ex->Xcanon, canon.size() * 2), "cudaMalloc(Xcanon)"))) return rc;
  if ((rc = check_cuda(cudaMalloc(&ex->gram, gram.size() * 8), "cudaMalloc(gram)"))) return rc;
  if ((rc = check_cuda(cudaMalloc(&ex->ref, REF_FLOATS * 4), "cudaMalloc(ref)"))) return rc;
  if ((rc = check_cuda(cudaMalloc(&ex->mean_acc, 32 * 8), "cudaMalloc(mean_acc)"))) return rc;
  if ((rc = check_cuda(cudaMalloc(&ex->qmean, 32 * 4), "cudaMalloc(qmean)"))) return rc;
  if ((rc = check_cuda(cudaMalloc(&ex->ident, 352 * 4), "cudaMalloc(ident)"))) return rc;
  if ((rc = check_cuda(cudaMalloc(&ex->zero, 4), "cudaMalloc(zero)"))) return rc;
  if ((rc = check_cuda(cudaMemcpy(ex->Xcanon, canon.data(), canon.size() * 2, cudaMemcpyHostToDevice), "cudaMemcpy"))) return rc;
  if ((rc = check_cuda(cudaMalloc(&ex->Xcanon64, canon64.size() * 2), "cudaMalloc(Xcanon64)"))) return rc;
  if ((rc = check_cuda(cudaMemcpy(ex->Xcanon64, canon64.data(), canon64.size() * 2, cudaMemcpyHostToDevice), "cudaMemcpy"))) return rc;
  if ((rc = check_cuda(cudaMemcpy(ex->gram, gram.data(), gram.size() * 8, cudaMemcpyHostToDevice), "cudaMemcpy"))) return rc;
  m->extra = ex;
  return AMCMC_OK;
}

void tc_release_l2(DiamondsTcExtra* ex, bool wait) {
  if (!ex->l2_dirty) return;
  if (ex->have_done_ev) {
    if (wait) cudaEventSynchronize(ex->done_ev);
    else if (cudaEventQuery(ex->done_ev) != cudaSuccess) { cudaGetLastError(); return; }  // still running: try again later
  }
  cudaCtxResetPersistingL2Cache();
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ex->l2_prev_limit);
  cudaGetLastError();
  ex->l2_dirty = 0;
}

void tc_run_begin(DiamondsTcExtra* ex, cudaStream_t s) {
  tc_release_l2(ex, false);
  if (ex->have_done_ev && ex->last_stream != s) cudaStreamWaitEvent(s, ex->done_ev, 0);
}

void tc_run_end(DiamondsTcExtra* ex, cudaStream_t s) {
  if (!ex->have_done_ev) {
    if (cudaEventCreateWithFlags(&ex->done_ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
    ex->have_done_ev = 1;
  }
  cudaEventRecord(ex->done_ev, s);
  ex->last_stream = s;
}

void destroy_diamonds_tc(amcmc_model* m) {
  DiamondsTcExtra* ex = (DiamondsTcExtra*)m->extra;
  if (!ex) return;
  tc_release_l2(ex, true);
  if (ex->have_done_ev) { cudaEventSynchronize(ex->done_ev); cudaEventDestroy(ex->done_ev); }
  if (ex->Xcanon) cudaFree(ex->Xcanon);
  if (ex->Xcanon64) cudaFree(ex->Xcanon64);
  if (ex->gram) cudaFree(ex->gram);
  if (ex->ref) cudaFree(ex->ref);
  if (ex->xprop) cudaFree(ex->xprop);
  if (ex->mean_acc) cudaFree(ex->mean_acc);
  if (ex->qmean) cudaFree(ex->qmean);
  if (ex->ident) cudaFree(ex->ident);
  if (ex->zero) cudaFree(ex->zero);
  if (ex->ldl) cudaFree(ex->ldl);
  if (ex->cref) cudaFree(ex->cref);
  if (ex->crss) cudaFree(ex->crss);
  free(ex);
  m->extra = nullptr;
}

bool diamonds_tc_available(const amcmc_model* m) { return m->model_id == AMCMC_MODEL_DIAMONDS && m->extra != nullptr; }

// common launch preparation: parking buffer, parameter block
int diamonds_tc_prepare(const amcmc_model* m, int64_t n_chains, TcParams* pp, const amcmc_state* st, const amcmc_run_args* a) {
  DiamondsTcExtra* ex = (DiamondsTcExtra*)m->extra;
  int rc;
  if (ex->xprop_cap < n_chains) {
    if (ex->xprop) cudaFree(ex->xprop);
    ex->xprop = nullptr;
    ex->xprop_cap = 0;
    if ((rc = check_cuda(cudaMalloc(&ex->xprop, (size_t)n_chains * TC_D * 4), "cudaMalloc(xprop)"))) return rc;
    ex->xprop_cap = n_chains;
  }
  TcParams& p = *pp;
  p.C = st->n_chains;
  p.n_groups = (int)((st->n_chains + TC_M - 1) / TC_M);
  p.Xcanon = ex->Xcanon;
  p.n_tiles = ex->n_tiles;
  p.ref = ex->ref;
  p.z = (float*)st->z;
  p.pe = (float*)st->potential_energy;
  p.macc = (float*)st->mean_accept_prob;
  p.xprop = ex->xprop;
  p.i0 = st->i;
  p.n_steps = a->n_steps;
  p.thinning = a->thinning;
  p.collect_start = a->collect_start;
  p.seed = a->seed;
  p.chain_offset = a->chain_offset;
  p.normals = (const float*)a->normals;
  p.uniforms = (const float*)a->uniforms;
  p.out_z = (float*)a->out_z;
  p.out_pe = (float*)a->out_potential_energy;
  p.out_acc = a->out_accept;
  p.n_rows = (float)m->n_rows;
  p.cst = m->cst;
  return AMCMC_OK;
}

void diamonds_tc_launch_ref(const amcmc_model* m, const float* loc, const float* scale, const float* lam, double eps,
                            cudaStream_t s) {
  DiamondsTcExtra* ex = (DiamondsTcExtra*)m->extra;
  diamonds_tc_ref_kernel<float><<<1, 128, 0, s>>>(ex->gram, loc, scale, lam, eps, ex->ref);
}

// Frozen / pooled run on the tensor cores: all chains share (loc, scale, log_step_size) given as
// device arrays of the model dtype (fp32).
int run_diamonds_tc(const amcmc_model* m, const amcmc_state* st, const void* loc, const void* scale_packed,
                    const void* log_step, const amcmc_run_args* a, cudaStream_t s) {
  DiamondsTcExtra* ex = (DiamondsTcExtra*)m->extra;
  if (!ex) { set_error("diamonds tensor-core path needs K = 25 predictors and fp32"); return AMCMC_ERR_UNSUPPORTED; }
  TcParams p;
  tc_run_begin(ex, s);
  int rc = diamonds_tc_prepare(m, st->n_chains, &p, st, a);
  if (rc) return rc;
  diamonds_tc_launch_ref(m, (const float*)loc, (const float*)scale_packed, (const float*)log_step, a->eps, s);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.n_groups < sms ? p.n_groups : sms;
  const bool ext = a->rng_mode == AMCMC_RNG_EXTERNAL;
  if (ext) {
    if ((rc = check_cuda(cudaFuncSetAttribute(diamonds_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::BYTES), "cudaFuncSetAttribute"))) return rc;
    diamonds_tc_kernel<true><<<grid, TC_THREADS, TcSmem::BYTES, s>>>(p);
  } else {
    if ((rc = check_cuda(cudaFuncSetAttribute(diamonds_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::BYTES), "cudaFuncSetAttribute"))) return rc;
    diamonds_tc_kernel<false><<<grid, TC_THREADS, TcSmem::BYTES, s>>>(p);
  }
  rc = check_cuda(cudaGetLastError(), "diamonds_tc_kernel launch");
  tc_run_end(ex, s);
  return rc;
}

}  // namespace amcmc
