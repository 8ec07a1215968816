// mmd_tc.cu -- the three Gaussian-kernel sums behind `mmd2_unbiased` / `mmd_heuristic`
// (reference: python/utils/evaluation.py:201-222 `gaussian_kernel`, :246-263, :279-294) as ONE tcgen05 launch:
//     sum_{i != j} k(x_i, x_j),   sum_{i != j} k(y_i, y_j),   sum_{ij} k(x_i, y_j),      k(a, b) = exp(-gamma |a - b|^2)
//
// |a - b|^2 = |a|^2 + |b|^2 - 2 a.b: the cross terms are a GEMM, the rest is an epilogue.  Both samples are
// centred on the mean of y (distances do not change, norms shrink) and split into two bfloat16 halves, a = hi + lo; the
// operands are concatenated along K as A' = [hi | lo | hi], B' = [hi | hi | lo], so one K' = 3d GEMM in bf16 with fp32
// accumulation gives hi.hi + lo.hi + hi.lo -- the product to ~2^-17 relative (the lo.lo term is dropped), zero-mean rounding.
// d <= 32 (K' <= 96): the reference evaluates d = 10 (eight_schools), 26 (diamonds), 4 (kidiq).
//
// Kernel: one persistent CTA per SM walks (A tile = 128 points, B tile = 256 points) pairs of the three problems.
//   warp 0     TMA producer: cp.async.bulk of the canonical K-major A tile and its scaled norms into a 2-stage ring; the B tile
//              (and its norms) stays in one of two buffers while the A tiles sweep past it (20 KB instead of 60 KB per pair)
//   warp 1     MMA issuer: K'/16 tcgen05.mma (M = 128, N = 256) into one of two 256-column TMEM accumulators
//   warps 2-17 epilogue: tcgen05.ld 32 columns at a time (the next chunk in flight while this one is consumed),
//              ex2.approx.ftz(fma(acc, 2 gamma log2e, na_i + nb_j)), fp32 partial sums per pair, float64 per thread over the launch
// An A stage (operand + norms) and its accumulator are released together by the epilogue, so the MMA of pair t + 1 overlaps the
// epilogue of pair t and the load of pair t + 2 starts when that epilogue ends.  The epilogue is bounded by MUFU (one ex2 per
// pair of points, 16 per clock per SM = 2,048 cycles per tile pair).  clock64 build (-DAMCMC_MMD_TIMING), n = m = 40,000:
// 3.15k cycles per tile pair = 226 waiting for the accumulator + 180 for the first TMEM load + 2,620 in the exp loop + 130
// hand-back: 0.65 of the MUFU bound, 0.78 inside the exp loop.  A 4-deep A ring with its own release barriers was built and
// measured: no faster (the loads were never exposed; the extra bookkeeping cost 10 %) -- what remains is that all 16 epilogue
// warps wait for the same accumulator at the same time; 128-column accumulators (four in flight) would let them stagger.
// The three sums at 10^4 points take 0.21 ms per call, launch overheads included, against 1.74 ms for the three CUDA-core
// passes of eval.cu.
#include <cmath>
#include <cstdio>
#include <cstring>
#include "internal.h"
#include "tc_common.cuh"

namespace amcmc {

using namespace tc;

constexpr int MT_M = 128;        // points per A tile (UMMA M)
constexpr int MT_N = 256;        // points per B tile (UMMA N)
constexpr int MT_KMAX = 96;      // 3 * 32
#ifndef AMCMC_MMD_EPI_WARPS
#define AMCMC_MMD_EPI_WARPS 16
#endif
constexpr int MT_EPI_WARPS = AMCMC_MMD_EPI_WARPS;  // 4 (2) per TMEM lane quarter: each drains 64 (128) of the 256 columns
constexpr int MT_EPI_COLS = MT_N / (MT_EPI_WARPS / 4);
constexpr int MT_THREADS = 32 * (2 + MT_EPI_WARPS);
constexpr int MT_A_BYTES = MT_M * MT_KMAX * 2;  // 24576 (upper bound; a launch moves M * K' * 2)
constexpr int MT_B_BYTES = MT_N * MT_KMAX * 2;  // 49152
constexpr int MT_STAGE = MT_A_BYTES + MT_M * 4;          // A operand + its norms: 25088
constexpr int MT_BBUF = MT_B_BYTES + MT_N * 4;           // B operand + its norms: 50176
constexpr int MT_OFF_B = 2 * MT_STAGE;
constexpr int MT_OFF_BAR = MT_OFF_B + 2 * MT_BBUF;
constexpr int MT_SMEM = MT_OFF_BAR + 128;

struct MmdParams {
  const uint16_t* tilesA[2];  // [sample] canonical 128-row tiles, [hi | lo | hi]
  const uint16_t* tilesB[2];  // [sample] canonical 256-row tiles, [hi | hi | lo]
  const float* normA[2];      // [sample] -gamma log2e |a|^2 per padded point (-inf for padding), 128 per tile
  const float* normB[2];      // the same, 256 per tile
  int nA[2], nB[2];           // tiles per sample
  int ksteps;                 // K' / 16
  float scale;                // 2 gamma log2e
  double* out;                // [3]
};

// centre, split and lay out one sample in both tile formats; one thread per (padded) point
__global__ void mmd_prep_kernel(const float* __restrict__ x, int64_t n, int d, const double* __restrict__ centre_sum, double inv_m,
                                float neg_gamma_log2e, int kp, uint16_t* __restrict__ tilesA, uint16_t* __restrict__ tilesB,
                                float* __restrict__ normA, float* __restrict__ normB, int64_t padA, int64_t padB) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float ninf = __int_as_float(0xff800000);
  if (i >= n) {
    if (i < padA) normA[i] = ninf;
    if (i < padB) normB[i] = ninf;
    return;
  }
  uint16_t* ta = tilesA + (i / MT_M) * (int64_t)(MT_M * kp);
  uint16_t* tb = tilesB + (i / MT_N) * (int64_t)(MT_N * kp);
  const int ra = (int)(i % MT_M), rb = (int)(i % MT_N);
  float nn = 0.f;
  for (int k = 0; k < d; ++k) {
    const float v = x[i * d + k] - (float)(centre_sum[k] * inv_m);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    const uint16_t hb = __bfloat16_as_ushort(h), lb = __bfloat16_as_ushort(l);
    nn = fmaf(v, v, nn);
    ta[canon_off(ra, k, MT_M) / 2] = hb;
    ta[canon_off(ra, d + k, MT_M) / 2] = lb;
    ta[canon_off(ra, 2 * d + k, MT_M) / 2] = hb;
    tb[canon_off(rb, k, MT_N) / 2] = hb;
    tb[canon_off(rb, d + k, MT_N) / 2] = hb;
    tb[canon_off(rb, 2 * d + k, MT_N) / 2] = lb;
  }
  normA[i] = neg_gamma_log2e * nn;
  normB[i] = neg_gamma_log2e * nn;
}

__global__ void mmd_colsum_kernel(const float* __restrict__ y, int64_t m, int d, double* __restrict__ out) {
  const int k = blockIdx.y;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) s += (double)y[i * d + k];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&out[k], s);
}

// 2^x, one MUFU.EX2 (results below 2^-126 flush to zero: they do not matter in a sum of kernel values)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// pair index -> problem (0: x.x, 1: y.y, 2: x.y), A tile, B tile; A tiles vary fastest so a CTA's consecutive pairs share B in L2
struct PairIndex {
  int prob, ta, tb, sa, sb;
};
__device__ __forceinline__ PairIndex decode_pair(const MmdParams& p, int64_t q) {
  PairIndex r;
  const int64_t p0 = (int64_t)p.nA[0] * p.nB[0], p1 = (int64_t)p.nA[1] * p.nB[1];
  if (q < p0) { r.prob = 0; r.sa = 0; r.sb = 0; }
  else if (q < p0 + p1) { r.prob = 1; r.sa = 1; r.sb = 1; q -= p0; }
  else { r.prob = 2; r.sa = 0; r.sb = 1; q -= p0 + p1; }
  r.tb = (int)(q / p.nA[r.sa]);
  r.ta = (int)(q % p.nA[r.sa]);
  return r;
}

__global__ void __launch_bounds__(MT_THREADS, 1) mmd_tc_kernel(const MmdParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MT_OFF_BAR);
  uint64_t* full = bars;           // [2] TMA landed
  uint64_t* acc_full = bars + 2;   // [2] MMA done
  uint64_t* acc_empty = bars + 4;  // [2] epilogue done: stage and accumulator free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], MT_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t total = (int64_t)p.nA[0] * p.nB[0] + (int64_t)p.nA[1] * p.nB[1] + (int64_t)p.nA[0] * p.nB[1];
  const int64_t per = (total + gridDim.x - 1) / gridDim.x;
  const int64_t q_begin = (int64_t)blockIdx.x * per;
  const int64_t q_end = q_begin + per < total ? q_begin + per : total;
  const uint32_t a_bytes = (uint32_t)(MT_M * p.ksteps * 32), b_bytes = (uint32_t)(MT_N * p.ksteps * 32);

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, cur_b = -1, bbuf = 1;
      for (int64_t q = q_begin; q < q_end; ++q, ++it) {
        const int s = it & 1;
        // pair it - 2 is through its epilogue: stage s, accumulator s AND the B buffer of two tiles ago are free (pairs finish
        // in order, and the previous B tile served at least pair it - 1)
        mbar_wait(&acc_empty[s], (uint32_t)(((it >> 1) & 1) ^ 1));
        const PairIndex pi = decode_pair(p, q);
        unsigned char* st = smem + s * MT_STAGE;
        const int bkey = pi.sb * (1 << 24) + pi.tb;
        const bool new_b = bkey != cur_b;  // a new B tile: into the other buffer, on this pair's barrier
        if (new_b) { cur_b = bkey; bbuf ^= 1; }
        unsigned char* sb = smem + MT_OFF_B + bbuf * MT_BBUF;
        mbar_arrive_expect_tx(&full[s], a_bytes + MT_M * 4 + (new_b ? b_bytes + MT_N * 4 : 0u));
        tma_load_1d(st, p.tilesA[pi.sa] + (int64_t)pi.ta * (a_bytes / 2), a_bytes, &full[s]);
        tma_load_1d(st + MT_A_BYTES, p.normA[pi.sa] + (int64_t)pi.ta * MT_M, MT_M * 4, &full[s]);
        if (new_b) {
          tma_load_1d(sb, p.tilesB[pi.sb] + (int64_t)pi.tb * (b_bytes / 2), b_bytes, &full[s]);
          tma_load_1d(sb + MT_B_BYTES, p.normB[pi.sb] + (int64_t)pi.tb * MT_N, MT_N * 4, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(MT_M, MT_N);
      constexpr uint32_t a_kstride = (MT_M / 8) * 128, b_kstride = (MT_N / 8) * 128;
      constexpr uint32_t kAChunk = (2 * a_kstride) >> 4, kBChunk = (2 * b_kstride) >> 4;
      int it = 0, cur_b = -1, bbuf = 1;
      for (int64_t q = q_begin; q < q_end; ++q, ++it) {
        const int s = it & 1;
        const PairIndex pi = decode_pair(p, q);
        const int bkey = pi.sb * (1 << 24) + pi.tb;
        if (bkey != cur_b) { cur_b = bkey; bbuf ^= 1; }
        unsigned char* st = smem + s * MT_STAGE;
        const uint64_t da = make_smem_desc(smem_u32(st), a_kstride, 128),
                       db = make_smem_desc(smem_u32(smem + MT_OFF_B + bbuf * MT_BBUF), b_kstride, 128);
        const uint32_t a_lo = (uint32_t)da, a_hi = (uint32_t)(da >> 32), b_lo = (uint32_t)db, b_hi = (uint32_t)(db >> 32);
        const uint32_t dcol = tmem_base + (uint32_t)(s * MT_N);
        mbar_wait(&full[s], (uint32_t)((it >> 1) & 1));  // the stage was released by the epilogue, which also freed accumulator s
        tc_fence_after();
        umma_bf16_lean<false>(dcol, a_lo, a_hi, b_lo, b_hi, idesc);
        for (int ks = 1; ks < p.ksteps; ++ks) umma_bf16_lean<true>(dcol, a_lo + ks * kAChunk, a_hi, b_lo + ks * kBChunk, b_hi, idesc);
        umma_commit(&acc_full[s]);
      }
    }
  } else {
    const int quarter = warp & 3;             // TMEM lanes this warp may read
    const int part_col = ((warp - 2) >> 2) * MT_EPI_COLS;  // first of this warp's columns
    const int row = quarter * 32 + lane;
    double tot0 = 0.0, tot1 = 0.0, tot2 = 0.0;
#ifdef AMCMC_MMD_TIMING
    long long tk[4] = {0, 0, 0, 0};
#endif
    int it = 0, cur_b = -1, bbuf = 1;
    for (int64_t q = q_begin; q < q_end; ++q, ++it) {
      const int s = it & 1;
      const PairIndex pi = decode_pair(p, q);
      const int bkey = pi.sb * (1 << 24) + pi.tb;
      if (bkey != cur_b) { cur_b = bkey; bbuf ^= 1; }
      const float* nA = reinterpret_cast<const float*>(smem + s * MT_STAGE + MT_A_BYTES);
      const float* nB = reinterpret_cast<const float*>(smem + MT_OFF_B + bbuf * MT_BBUF + MT_B_BYTES) + part_col;
#ifdef AMCMC_MMD_TIMING  // clock64 breakdown of the epilogue (profiles/r02_eval.md)
      long long tl = clock64();
#define MT_T(k) { const long long c_ = clock64(); tk[k] += c_ - tl; tl = c_; }
#else
#define MT_T(k)
#endif
      mbar_wait(&full[s], (uint32_t)((it >> 1) & 1));  // the norms came with the same TMA transaction (already complete here)
      mbar_wait(&acc_full[s], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      MT_T(0)
      const float na = nA[row];
      const int64_t gi = (int64_t)pi.ta * MT_M + row, gj0 = (int64_t)pi.tb * MT_N + part_col;
      // same-sample problems skip i == j: only tile pairs that touch the diagonal take the compare
      const bool diag = pi.prob < 2 && gi >= gj0 && gi < gj0 + MT_EPI_COLS;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * MT_N + part_col);
      float part = 0.f;
      float v[2][32];
      tmem_ld_32x32(tbase, v[0]);
#pragma unroll
      for (int ch = 0; ch < MT_EPI_COLS / 32; ++ch) {
        tmem_ld_wait();
        MT_T(1)
        if (ch + 1 < MT_EPI_COLS / 32) tmem_ld_32x32(tbase + (uint32_t)((ch + 1) * 32), v[(ch + 1) & 1]);
        const float(&w)[32] = v[ch & 1];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 nb = *reinterpret_cast<const float4*>(nB + ch * 32 + c);
          float e0 = ex2_ftz(fmaf(w[c], p.scale, na + nb.x));
          float e1 = ex2_ftz(fmaf(w[c + 1], p.scale, na + nb.y));
          float e2 = ex2_ftz(fmaf(w[c + 2], p.scale, na + nb.z));
          float e3 = ex2_ftz(fmaf(w[c + 3], p.scale, na + nb.w));
          if (diag) {
            const int64_t gj = gj0 + ch * 32 + c;
            e0 = gi == gj ? 0.f : e0;
            e1 = gi == gj + 1 ? 0.f : e1;
            e2 = gi == gj + 2 ? 0.f : e2;
            e3 = gi == gj + 3 ? 0.f : e3;
          }
          part += (e0 + e1) + (e2 + e3);
        }
        MT_T(2)
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[s]);
      tot0 += pi.prob == 0 ? (double)part : 0.0;
      tot1 += pi.prob == 1 ? (double)part : 0.0;
      tot2 += pi.prob == 2 ? (double)part : 0.0;
      MT_T(3)
    }
#ifdef AMCMC_MMD_TIMING
    if (blockIdx.x == 0 && tid == 64 && it > 0)
      printf("[mmd] %d pairs: wait acc_full %lld, wait tmem ld %lld, exp %lld, tail %lld cycles per pair\n", it, tk[0] / it, tk[1] / it,
             tk[2] / it, tk[3] / it);
#endif
    for (int o = 16; o; o >>= 1) {
      tot0 += __shfl_xor_sync(0xffffffffu, tot0, o);
      tot1 += __shfl_xor_sync(0xffffffffu, tot1, o);
      tot2 += __shfl_xor_sync(0xffffffffu, tot2, o);
    }
    if (lane == 0) {
      if (tot0 != 0.0) atomicAdd(&p.out[0], tot0);
      if (tot1 != 0.0) atomicAdd(&p.out[1], tot1);
      if (tot2 != 0.0) atomicAdd(&p.out[2], tot2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// per-device workspace that grows on demand (a cudaMalloc / cudaFree pair per call would cost more than the kernel)
struct MmdWorkspace {
  char* buf;
  size_t bytes;
};
static MmdWorkspace* mmd_workspace(size_t need) {
  static MmdWorkspace ws[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (ws[dev].bytes < need) {
    if (ws[dev].buf) cudaFree(ws[dev].buf);
    ws[dev].buf = nullptr;
    ws[dev].bytes = 0;
    if (cudaMalloc(&ws[dev].buf, need) != cudaSuccess) return nullptr;
    ws[dev].bytes = need;
  }
  return &ws[dev];
}

}  // namespace amcmc

using namespace amcmc;

// x: DEVICE float32 [n][d], y: DEVICE float32 [m][d] (row-major, d <= 32).  out_host[3] (float64):
//   sum_{i != j} k(x_i, x_j), sum_{i != j} k(y_i, y_j), sum_{ij} k(x_i, y_j)   with k(a, b) = exp(-gamma |a - b|^2).
// The diagonal terms of the same-sample sums are exactly 1 each: add n (m) for the biased estimate of mmd_heuristic.
extern "C" int amcmc_eval_mmd_sums(const float* x, int64_t n, const float* y, int64_t m, int d, double gamma, double* out_host,
                                   void* stream) {
  if (!x || !y || !out_host || n < 1 || m < 1 || d < 1 || !(gamma >= 0.0) || !std::isfinite(gamma)) {
    set_error("amcmc_eval_mmd_sums: bad argument");
    return AMCMC_ERR_ARG;
  }
  if (d > 32 || n > (1 << 24) || m > (1 << 24)) {
    set_error("amcmc_eval_mmd_sums: the tensor-core path takes d <= 32 (use amcmc_eval_kernel_sum)");
    return AMCMC_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int ksteps = (3 * d + 15) / 16, kp = 16 * ksteps;
  const int64_t cnt[2] = {n, m};
  int nA[2], nB[2];
  size_t off = 512, oA[2], oB[2], onA[2], onB[2];
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  for (int k = 0; k < 2; ++k) {
    nA[k] = (int)((cnt[k] + MT_M - 1) / MT_M);
    nB[k] = (int)((cnt[k] + MT_N - 1) / MT_N);
    oA[k] = off; off += al((size_t)nA[k] * MT_M * kp * 2);
    oB[k] = off; off += al((size_t)nB[k] * MT_N * kp * 2);
    onA[k] = off; off += al((size_t)nA[k] * MT_M * 4);
    onB[k] = off; off += al((size_t)nB[k] * MT_N * 4);
  }
  MmdWorkspace* ws = mmd_workspace(off);
  if (!ws) { set_error("amcmc_eval_mmd_sums: workspace allocation failed"); return AMCMC_ERR_CUDA; }
  char* buf = ws->buf;
  double* scal = (double*)buf;  // [0..2] sums, [3..3+d) column sums of y (512-byte header)
  int rc;
  if ((rc = check_cuda(cudaMemsetAsync(buf, 0, off, s), "cudaMemsetAsync(mmd workspace)"))) return rc;  // zero K padding and row padding
  {
    int gx = (int)((m + 255) / 256);
    if (gx > 148) gx = 148;
    mmd_colsum_kernel<<<dim3(gx, d), 256, 0, s>>>(y, m, d, scal + 3);
  }
  const double log2e = 1.4426950408889634;
  const float* src[2] = {x, y};
  MmdParams p;
  for (int k = 0; k < 2; ++k) {
    const int64_t padA = (int64_t)nA[k] * MT_M, padB = (int64_t)nB[k] * MT_N;
    const int64_t threads = padB > padA ? padB : padA;
    mmd_prep_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(src[k], cnt[k], d, scal + 3, 1.0 / (double)m, (float)(-gamma * log2e), kp,
                                                                      (uint16_t*)(buf + oA[k]), (uint16_t*)(buf + oB[k]), (float*)(buf + onA[k]),
                                                                      (float*)(buf + onB[k]), padA, padB);
    p.tilesA[k] = (const uint16_t*)(buf + oA[k]);
    p.tilesB[k] = (const uint16_t*)(buf + oB[k]);
    p.normA[k] = (const float*)(buf + onA[k]);
    p.normB[k] = (const float*)(buf + onB[k]);
    p.nA[k] = nA[k];
    p.nB[k] = nB[k];
  }
  p.ksteps = ksteps;
  p.scale = -2.0f * (float)(-gamma * log2e);  // exactly twice the (rounded) factor of the norms: arg = c (|a|^2 + |b|^2 - 2 a.b) with ONE c
  p.out = scal;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    if ((rc = check_cuda(cudaFuncSetAttribute(mmd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM), "cudaFuncSetAttribute(mmd_tc_kernel)")))
      return rc;
    attr_set[dev & 63] = true;
  }
  const int64_t total = (int64_t)nA[0] * nB[0] + (int64_t)nA[1] * nB[1] + (int64_t)nA[0] * nB[1];
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(total < sms ? total : sms);
  mmd_tc_kernel<<<grid, MT_THREADS, MT_SMEM, s>>>(p);
  if ((rc = check_cuda(cudaGetLastError(), "mmd_tc_kernel launch"))) return rc;
  if ((rc = check_cuda(cudaMemcpyAsync(out_host, scal, 3 * sizeof(double), cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync"))) return rc;
  return check_cuda(cudaStreamSynchronize(s), "mmd_tc_kernel");
}
