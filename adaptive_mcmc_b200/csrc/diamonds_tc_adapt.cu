// diamonds_tc_adapt.cu -- diamonds on the tensor cores WITH per-chain adaptation: the full ARWMH.sample of the
// reference (python/kernels/arwmh.py:140-207) for tens of thousands of chains, the N-row likelihood on tcgen05
// (same GEMM formulation, operand tiles, TMEM epilogue and warp roles as diamonds_tc.cu) and every chain's own
// running mean, LDL^T proposal factor and step size updated between the tensor-core phases.
//
// Per-chain adaptation state cannot stay on chip here (128 chains x 1.4 KB per UMMA tile row block, four
// blocks per SM), so it lives in global memory and is L2-resident (92 MB at 65,536 chains).  The running mean and the
// positions stay in the struct-of-arrays layout of the ABI (chain index fastest: the 128 owner threads of a group
// touch them with coalesced accesses).  The proposal factor is re-laid out for the launch: LDL^T form, one block
// of [351 entries][128 chains] per group, entries in column-major order (TcLdl below).  The strides are then
// compile-time constants, so every access of the unrolled pass is `base + immediate` and costs no address
// arithmetic; two small kernels convert from / to the packed Cholesky form of the ABI around the launch.
// At every step boundary the owner thread of a chain makes ONE pass over its factor:
//     rank-one update with delta = x_new - loc (Gill-Golub-Murray-Saunders recurrence, as arwmh_small.cuh)
//     + accumulation of the NEXT proposal  L~'(sqrt(D') .* z_next)  while the entry is in a register,
// so the factor is read once and written once per step.  The helper warps run one step ahead and only
// produce the draws (z, u).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>
#include "diamonds_tc.cuh"

namespace amcmc {

struct TcAdaptParams {
  TcParams p;        // chains, tiles, positions, energies, draws, outputs (as the shared-state kernel)
  float* loc;        // [26][C]
  float* ldl;        // [n_groups][351][128]  LDL^T blocks
  const float* cref; // [50][C] per-chain reference point of the centred GEMM: q_ref (25), 2 g (25)
  const double* crss;  // [C] RSS at the reference point
  const uint16_t* Xcanon64;  // design-matrix tiles in this kernel's K = 64 layout
  float* lam;        // [C]
  float* asc;        // [C]
  int64_t num_warmup;
  float lr_decay, target, eps;
  int rnd_begin, rnd_end;  // rounds (sets of TC_GR groups per CTA) served by this launch
  uint32_t x_bytes;        // bytes fetched per design-matrix tile (TCA_TILE_BYTES; less only in the traffic experiment)
  uint32_t dbg;            // timing experiments (results wrong): 2 = two MMAs per accumulator, 4 = no factor pass
};

__device__ __forceinline__ int tri_f(int i, int j) { return i * (i + 1) / 2 + j; }

constexpr int TC_NE = TC_D * (TC_D + 1) / 2;  // 351 packed entries
__host__ __device__ constexpr int cm_col(int e) { int j = 0, len = TC_D; while (e >= len) { e -= len; --len; ++j; } return j; }
__host__ __device__ constexpr int cm_row(int e) { int j = 0, len = TC_D; while (e >= len) { e -= len; --len; ++j; } return j + e; }

// packed Cholesky of the ABI ([351][C], row-major packed) -> LDL^T blocks ([group][351 column-major][128]):
// slot of (j,j) = D_j = L_jj^2, slot of (i,j) = L~_ij = L_ij / L_jj.  Lanes past C get the identity.
__global__ void tc_chol_to_ldl_kernel(const float* __restrict__ sc, float* __restrict__ ldl, int64_t C) {
  const int64_t c = (int64_t)blockIdx.x * TC_M + threadIdx.x;
  float* blk = ldl + (int64_t)blockIdx.x * TC_NE * TC_M + threadIdx.x;
  int e = 0;
#pragma unroll 1
  for (int j = 0; j < TC_D; ++j) {
    const float dg = c < C ? sc[(int64_t)tri_f(j, j) * C + c] : 1.f;
    const float inv = 1.0f / dg;
    blk[(e++) * TC_M] = dg * dg;
    for (int i = j + 1; i < TC_D; ++i) blk[(e++) * TC_M] = c < C ? sc[(int64_t)tri_f(i, j) * C + c] * inv : 0.f;
  }
}
__global__ void tc_ldl_to_chol_kernel(const float* __restrict__ ldl, float* __restrict__ sc, int64_t C) {
  const int64_t c = (int64_t)blockIdx.x * TC_M + threadIdx.x;
  if (c >= C) return;
  const float* blk = ldl + (int64_t)blockIdx.x * TC_NE * TC_M + threadIdx.x;
  int e = 0;
#pragma unroll 1
  for (int j = 0; j < TC_D; ++j) {
    const float sd = sqrtf(blk[(e++) * TC_M]);
    sc[(int64_t)tri_f(j, j) * C + c] = sd;
    for (int i = j + 1; i < TC_D; ++i) sc[(int64_t)tri_f(i, j) * C + c] = blk[(e++) * TC_M] * sd;
  }
}

// Per-chain reference point of the centred GEMM (see diamonds_tc.cu for the formulation): q_ref = the chain's own
// position at the start of the segment, g = X1^T r_ref = h - G q_ref and RSS_ref = yy - 2 h.q + q.G q from the float64
// Gram matrix.  With its own reference a chain only ever sends its displacement SINCE THE SEGMENT START through the
// split-bf16 GEMM, so the accuracy of the energies does not depend on how far the batch is spread (chains that start
// at U(-2,2)^26 sit at |U| ~ 1e4-1e6 for tens of thousands of steps; a shared reference costs 30x accuracy there).
__global__ void tc_chain_ref_kernel(const double* __restrict__ gram, const float* __restrict__ z, int64_t C, float* __restrict__ cref,
                                    double* __restrict__ crss) {
  __shared__ double sG[TC_KC * TC_KC + TC_KC + 1];
  for (int e = threadIdx.x; e < TC_KC * TC_KC + TC_KC + 1; e += blockDim.x) sG[e] = gram[e];
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* h = sG + TC_KC * TC_KC;
  double q[TC_KC];
#pragma unroll
  for (int k = 0; k < TC_KC; ++k) {
    const float v = z[(int64_t)k * C + c];
    q[k] = (double)v;
    cref[(int64_t)k * C + c] = v;
  }
  double rss = sG[TC_KC * TC_KC + TC_KC];
#pragma unroll 1
  for (int a = 0; a < TC_KC; ++a) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < TC_KC; ++k) s = fma(sG[a * TC_KC + k], q[k], s);
    cref[(int64_t)(TC_KC + a) * C + c] = (float)(2.0 * (h[a] - s));
    double qa = 0.0;  // q[a] with a runtime index: select instead of indexing (keeps q in registers)
#pragma unroll
    for (int k = 0; k < TC_KC; ++k) qa = (k == a) ? q[k] : qa;
    rss = fma(qa, s - 2.0 * h[a], rss);
  }
  crss[c] = rss;
}

// One pass over the chain's LDL^T factor in global memory (see the file header).
//   upd    : apply (1-gamma) L D L^T + gamma w w^T (w is consumed).  Without it the caller passes gamma = 0 and
//            w = 0, which makes every expression below reproduce the old factor bit for bit, and nothing is stored.
//   WANT   : accumulate |L' e^lam' - L e^lam|_F^2  (arwmh.py:197)
//   acc_i  = sum_{j<=i} L~'_ij sqrt(D'_j) zn_j     (zn: next draws in shared memory, stride TC_M; zero if !have_next)
// (Evict-first hints on the factor stream, ld.cs / st.cs, were measured and make the pass 1.6x slower.)
// The 351 entries are visited column by column through a ring of TC_RING registers: entry e + TC_RING is requested
// when entry e is consumed, so TC_RING loads stay in flight per thread for the whole pass (the pass is bound by L2
// latency, not by bandwidth or issue).  The visit is unrolled at compile time (fold over an index sequence) so the
// ring, w and acc are registers; a rolled loop would put them in local memory, which misses the ~30 KB of L1 left
// beside the tiles.
constexpr int TC_RING = 64;

template <bool WANT>
struct PassCtx {
  float* col;      // this chain's lane of its group block: entry e at col[e * TC_M]
  const float* zn;
  bool have_next, upd;
  float gamma, omg, el_old, el_new;
  float t, ss, wj, coef, yj, so, sn;
  float ring[TC_RING];
  float w[TC_D], acc[TC_D];
};

template <bool WANT, int E>
__device__ __forceinline__ void tc_pass_entry(PassCtx<WANT>& x) {
  constexpr int j = cm_col(E), i = cm_row(E);
  const float v = x.ring[E % TC_RING];
  if constexpr (E + TC_RING < TC_NE) x.ring[E % TC_RING] = x.col[(E + TC_RING) * TC_M];
  if constexpr (i == j) {
    const float Dold = v;
    const bool pos = Dold > 0.f;  // a non-positive pivot leaves its column untouched
    const float Dj = x.omg * Dold;
    const float wj = x.w[j];
    const float cw = x.gamma * wj;
    const float g = fmaf(cw * wj, x.t, Dj);
    const float tr = __fdividef(x.t, g);
    const float Dnew = pos ? g : Dold;
    x.coef = pos ? cw * tr : 0.f;
    x.t = pos ? Dj * tr : x.t;
    x.wj = pos ? wj : 0.f;
    if (x.upd) x.col[E * TC_M] = Dnew;
    const float sn_raw = sqrtf(Dnew);
    x.yj = x.have_next ? sn_raw * x.zn[j * TC_M] : 0.f;
    x.acc[j] += x.yj;
    if (WANT) {
      x.so = sqrtf(Dold) * x.el_old;
      x.sn = sn_raw * x.el_new;
      const float dd = x.sn - x.so;
      x.ss = fmaf(dd, dd, x.ss);
    }
  } else {
    const float Lo = v;
    x.w[i] = fmaf(-x.wj, Lo, x.w[i]);
    const float Ln = fmaf(x.coef, x.w[i], Lo);
    if (x.upd) x.col[E * TC_M] = Ln;
    x.acc[i] = fmaf(Ln, x.yj, x.acc[i]);
    if (WANT) {
      const float df = fmaf(Ln, x.sn, -(Lo * x.so));
      x.ss = fmaf(df, df, x.ss);
    }
  }
}
template <bool WANT, size_t... E>
__device__ __forceinline__ void tc_pass_all(PassCtx<WANT>& x, std::index_sequence<E...>) {
  (tc_pass_entry<WANT, (int)E>(x), ...);
}
template <bool WANT, size_t... E>
__device__ __forceinline__ void tc_pass_fill(PassCtx<WANT>& x, std::index_sequence<E...>) {
  ((x.ring[E] = x.col[(int)E * TC_M]), ...);
}

template <bool WANT>
__device__ __forceinline__ float tc_column_pass(float* __restrict__ col, const float* zn, bool have_next,
                                                bool upd, float (&w)[TC_D], float gamma, float el_old, float el_new,
                                                float (&acc)[TC_D]) {
  PassCtx<WANT> x;
  x.col = col; x.zn = zn; x.have_next = have_next; x.upd = upd;
  x.gamma = gamma; x.omg = 1.f - gamma; x.el_old = el_old; x.el_new = el_new;
  x.t = 1.f; x.ss = 0.f; x.wj = 0.f; x.coef = 0.f; x.yj = 0.f; x.so = 0.f; x.sn = 0.f;
  tc_pass_fill<WANT>(x, std::make_index_sequence<TC_RING>{});
#pragma unroll
  for (int k = 0; k < TC_D; ++k) { x.w[k] = w[k]; x.acc[k] = 0.f; }
  tc_pass_all<WANT>(x, std::make_index_sequence<TC_NE>{});
#pragma unroll
  for (int k = 0; k < TC_D; ++k) acc[k] = x.acc[k];
  return x.ss;
}

// Operand layout: the K = 80 concatenation of diamonds_tc.cu -- A' = [D_hi | D_lo | D_hi | 0] (25 + 25 + 25 + 5 columns),
// B' = [X_hi | X_hi | X_lo | 0], five K = 16 MMAs per accumulator; the design-matrix tiles are the ones the shared-state
// kernel uses.  Round 1 used K = 64 operands ([D_hi (25 + 7 zeros) | D_lo (25 + 7 zeros)] x [X_hi | X_lo], six MMAs whose
// descriptors pick the chunk pairs; -DAMCMC_TCA_K64): 20 % fewer bytes per tile, which mattered when the kernel was thought to
// be short of SM <-> L2 bandwidth.  It is not (profiles/r02_diamonds_tc_adaptive.md): it runs at the board's power limit, where
// one MMA less per accumulator is worth more -- 33.2 vs 34.4 ms per 500 steps at 65,536 chains.
#ifndef AMCMC_TCA_K64
constexpr int TCA_K = 80;
#else
constexpr int TCA_K = 64;
#endif
constexpr int TCA_TILE_BYTES = TC_TILE_N * TCA_K * 2;  // 32768
constexpr int TCA_A_BYTES = TC_M * TCA_K * 2;          // 16384
constexpr int TCA_STAGES = 2;                          // design-matrix tiles in flight (a third stage was measured: no gain)
// Accumulator width (UMMA N): 256 data rows = one design-matrix tile per accumulator, two accumulators in the 512 TMEM
// columns, one per MMA issuer.  (-DAMCMC_TCA_ACC_N=128 gives four half-tile accumulators; measured slower, 119 k vs 104 k
// cycles per step at 65,536 chains: the hand-off cost is per accumulator, not per column.)
#ifndef AMCMC_TCA_ACC_N
#define AMCMC_TCA_ACC_N 256
#endif
constexpr int TCA_ACC_N = AMCMC_TCA_ACC_N;             // 128 (four accumulators in flight) or 256 (two)
constexpr int TCA_NBUF = 512 / TCA_ACC_N;
constexpr int TCA_HALVES = TC_TILE_N / TCA_ACC_N;      // accumulators per design-matrix tile and group
struct TcaSmem {
  static constexpr int OFF_X = 0;                                // TCA_STAGES stages
  static constexpr int OFF_A = TCA_STAGES * TCA_TILE_BYTES;      // TC_GR groups
  static constexpr int OFF_BAR = OFF_A + TC_GR * TCA_A_BYTES;    // x_full[2], x_empty[2], (4 unused), a_ready, v_full, v_empty
  static constexpr int OFF_TMEM = OFF_BAR + TcSmem::N_BAR * 8;
  static constexpr int OFF_ACCBAR = OFF_TMEM + 16;               // acc_full[2][TCA_NBUF], acc_empty[2][TCA_NBUF]
  static constexpr int OFF_V = OFF_ACCBAR + 2 * 2 * TCA_NBUF * 8;  // float [TC_GR][27][TC_M]
  static constexpr int BYTES = OFF_V + TC_GR * 27 * TC_M * 4;
};

// proposal -> A' row (split bf16), scalar part of U', shadow position buffer.  Returns via references.
__device__ __forceinline__ void tc_emit_proposal(const float (&xp)[TC_D], const float* __restrict__ cref, int64_t C, double rss_ref,
                                                 unsigned char* sA, int g, int row, uint64_t* a_ready_g, double n_rows, double cst,
                                                 double& Up_part, double& inv2var, double& rss_part) {
  float dq = 0.f;
  uint32_t hi[TC_KC], lo[TC_KC];  // bf16 bit patterns in the low halves
  float qr[TC_KC], g2[TC_KC];
#pragma unroll
  for (int k = 0; k < TC_KC; ++k) {  // this chain's reference point and gradient (coalesced over the group)
    qr[k] = __ldg(cref + (int64_t)k * C);
    g2[k] = __ldg(cref + (int64_t)(TC_KC + k) * C);
  }
#pragma unroll
  for (int k = 0; k < TC_KC; ++k) {
    const float dlt = xp[k] - qr[k];
    dq = fmaf(dlt, g2[k], dq);
    const uint16_t h = f2bf(dlt);
    hi[k] = h;
    lo[k] = f2bf(dlt - bf2f(h));
  }
#ifndef AMCMC_TCA_K64
  // K layout: [hi(25) | lo(25) | hi(25) | 0(5)]
  auto elem = [&](int k) -> uint32_t { return k < TC_KC ? hi[k] : (k < 2 * TC_KC ? lo[k - TC_KC] : (k < 3 * TC_KC ? hi[k - 2 * TC_KC] : 0u)); };
#else
  // K layout: [hi(25) | 0(7) | lo(25) | 0(7)]
  auto elem = [&](int k) -> uint32_t { return k < TC_KC ? hi[k] : (k < 32 ? 0u : (k < 32 + TC_KC ? lo[k - 32] : 0u)); };
#endif
  unsigned char* arow = sA + g * TCA_A_BYTES + (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
  for (int kc = 0; kc < TCA_K / 8; ++kc) {
    uint4 v;
    v.x = elem(8 * kc) | (elem(8 * kc + 1) << 16);
    v.y = elem(8 * kc + 2) | (elem(8 * kc + 3) << 16);
    v.z = elem(8 * kc + 4) | (elem(8 * kc + 5) << 16);
    v.w = elem(8 * kc + 6) | (elem(8 * kc + 7) << 16);
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
            (n_rows - 1.0) * (double)s + cst;
  rss_part = rss_ref - (double)dq;  // + sum m^2 from the GEMM = RSS; assembled and clamped at 0 by the caller
}

// Warp roles: warps 0-7 = the two sampler warpgroups (TMEM epilogue + per-chain state), warp 8 = TMA producer,
// warp 9 = MMA issuer, warps 10-11 = draws.  The sampler warpgroups take the registers the third one gives up
// (setmaxnreg: 2 x 128 x 208 + 128 x 88 = 384 x 168, the CTA's pool at launch), which is what keeps the unrolled column pass spill-free.
constexpr int64_t kSegment = 256;  // steps between moves of the GEMM reference point
// Warp roles of the service warpgroup.  TWO MMA issuers: tcgen05.mma issue blocks the issuing thread while the tensor pipe is
// busy (its queue is one deep: scripts/probes/umma_bench.cu), so with a single issuer every barrier wait, fence and commit
// between two accumulators idled the tensor pipe -- measured 1,525 cycles per 256-row accumulator for 768 cycles of tensor
// work.  Issuer A serves the first group of the active stream (TMEM buffers 0 .. NBI-1), issuer B the second
// (buffers NBI ..): while one waits for its buffer to be drained the other's MMAs run.  The draws come from one warp.
constexpr int kTmaWarp = TC_EPI_WARPS, kMmaWarp = TC_EPI_WARPS + 1, kDrawWarp = TC_EPI_WARPS + 2, kMmaWarpB = TC_EPI_WARPS + 3;
constexpr int TCA_NBI = TCA_NBUF / 2;  // TMEM buffers per issuer
constexpr int kSamplerRegs = 208, kServiceRegs = 88, kLaunchRegs = 168;  // launch: 65536 / 384 rounded down to 8
static_assert(TC_EPI_WARPS == 8 && TC_THREADS == 384, "warp roles assume 8 sampler + 4 service warps");
static_assert(32 * TC_EPI_WARPS * kSamplerRegs + 128 * kServiceRegs <= TC_THREADS * kLaunchRegs, "the pool is what the CTA got at launch");

#ifdef AMCMC_TC_SPIN
#define TC_HOT_WAIT mbar_wait_spin
#else
#define TC_HOT_WAIT mbar_wait
#endif
#ifdef AMCMC_TC_TIMING
__device__ unsigned long long tc_dbg[64];
// phase clocks go to shared memory (a global read-modify-write per sample would itself cost ~1000 cycles) and are flushed once
#define TC_T(k) do { if (dbg) { const long long now_ = clock64(); s_dbg[k] += (unsigned long long)(now_ - tlast); tlast = now_; } } while (0)
#define TC_W(k) do { if (dbg) { const long long now_ = clock64(); s_dbg_w[k] += (unsigned long long)(now_ - tlast); tlast = now_; } } while (0)
#else
#define TC_T(k) do { } while (0)
#define TC_W(k) do { } while (0)
#endif

template <bool EXTERNAL>
__global__ void __launch_bounds__(TC_THREADS, 1) diamonds_tc_adapt_kernel(const TcAdaptParams ap) {
  const TcParams& p = ap.p;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sX = smem + TcaSmem::OFF_X;
  unsigned char* sA = smem + TcaSmem::OFF_A;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TcaSmem::OFF_BAR);
  uint64_t* x_full = bars;          // [TCA_STAGES] (the slots of the unused acc barriers of the shared layout follow)
  uint64_t* x_empty = bars + 4;     // [TCA_STAGES]
  // accumulator hand-off, per stream and TMEM buffer: [stream * TCA_NBUF + buffer].  Both streams use all buffers; a
  // waiter on an mbarrier parity may lag at most one phase, so the streams cannot share one barrier ring.
  uint64_t* acc_full = reinterpret_cast<uint64_t*>(smem + TcaSmem::OFF_ACCBAR);
  uint64_t* acc_empty = acc_full + 2 * TCA_NBUF;
  uint64_t* a_ready = bars + 8;
  uint64_t* v_full = bars + 8 + TC_GR;
  uint64_t* v_empty = bars + 8 + 2 * TC_GR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TcaSmem::OFF_TMEM);
  float* sV = reinterpret_cast<float*>(smem + TcaSmem::OFF_V);  // [TC_GR][27][TC_M]: next draws z[26], u

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // tells the compiler the role branches are warp-uniform
#ifdef AMCMC_TC_TIMING
  __shared__ unsigned long long s_dbg[64];
  if (tid < 64) s_dbg[tid] = 0;
#endif
  // Group ownership is interleaved: CTA b owns groups b, b + grid, b + 2 grid, ...  The first two groups of every CTA
  // are then the first 2 * grid groups of the factor buffer -- the contiguous prefix the host pins in L2 -- and both
  // streams of a CTA (groups 0, 2 / groups 1, 3) get one pinned and one streaming group each.
  const int g_count = (int)blockIdx.x < p.n_groups ? (p.n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < TCA_STAGES; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 2);  // one commit from each MMA issuer
    }
    for (int s = 0; s < 2 * TCA_NBUF; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], TC_EPI_WARPS / 2);  // one arrival from each of the four sampler warps of the owning stream
    }
    for (int g = 0; g < TC_GR; ++g) {
      mbar_init(&a_ready[g], TC_M);
      mbar_init(&v_full[g], 32);  // the draw warp
      mbar_init(&v_empty[g], TC_M);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_rounds_all = (g_count + TC_GR - 1) / TC_GR;
  const int n_rounds = n_rounds_all < ap.rnd_end ? n_rounds_all : ap.rnd_end;
  uint32_t x_it = 0, acc_it = 0, acc_it1 = 0, a_it = 0;  // acc_it / acc_it1: accumulators drained of the stream's first / second group
  uint32_t ks0 = 0, ks1 = 0;  // MMA warp: accumulators issued so far per stream

#define TC_ROUND_BEGIN                                   \
  for (int rnd = ap.rnd_begin; rnd < n_rounds; ++rnd) {  \
    const int G = min(TC_GR, g_count - rnd * TC_GR);     \
    const int64_t g0 = (int64_t)blockIdx.x + (int64_t)rnd * TC_GR * gridDim.x; \
    const int64_t gs = gridDim.x;                        \
    (void)g0; (void)gs;
#define TC_ROUND_END                                     \
    a_it += (uint32_t)p.n_steps;                         \
  }
#define TC_SERVICE_REGS asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kServiceRegs))

  {
    if (warp >= TC_EPI_WARPS) TC_SERVICE_REGS;  // one instruction for the whole service warpgroup (.aligned)
    if (warp == kTmaWarp) {
      TC_ROUND_BEGIN
      if (lane == 0) {  // ===== TMA producer =====
        const int n_streams = G > 1 ? 2 : 1;  // one sweep over the design matrix per stream and step
        for (int64_t sweep = 0; sweep < p.n_steps * n_streams; ++sweep)
          for (int tile = 0; tile < p.n_tiles; ++tile, ++x_it) {
            const int s = x_it % TCA_STAGES;
            mbar_wait(&x_empty[s], ((x_it / TCA_STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(&x_full[s], ap.x_bytes);
            tma_load_1d(sX + s * TCA_TILE_BYTES, ap.Xcanon64 + (size_t)tile * (TCA_TILE_BYTES / 2), ap.x_bytes, &x_full[s]);
          }
      }
      TC_ROUND_END
    } else if (warp == kMmaWarp || warp == kMmaWarpB) {
      const int issuer = warp == kMmaWarp ? 0 : 1;
      TC_ROUND_BEGIN
      if (lane == 0) {  // ===== MMA issuers =====
#ifdef AMCMC_TC_TIMING
        const bool dbg = (blockIdx.x == 0 && issuer == 0);
        long long tlast = clock64();
#endif
        const uint32_t idesc = make_idesc_bf16_f32(TC_M, TCA_ACC_N);
        constexpr uint32_t a_kstride = (TC_M / 8) * 128, b_kstride = (TC_TILE_N / 8) * 128;
        // base descriptors (group 0 / stage 0 / rows 0..): the loops below only add chunk offsets to the low words
        const uint64_t da0 = make_smem_desc(smem_u32(sA), a_kstride, 128), db0 = make_smem_desc(smem_u32(sX), b_kstride, 128);
        const uint32_t da_lo0 = (uint32_t)da0, da_hi = (uint32_t)(da0 >> 32), db_lo0 = (uint32_t)db0, db_hi = (uint32_t)(db0 >> 32);
        constexpr uint32_t kAChunk = (2 * a_kstride) >> 4, kBChunk = (2 * b_kstride) >> 4;  // one K = 16 chunk, in descriptor units
        for (int64_t st = 0; st < p.n_steps; ++st)
        for (int strm = 0; strm < 2; ++strm) {  // stream = groups strm, strm + 2 (see the sampler warps)
          if (strm >= G) break;
          const int g = strm + 2 * issuer;      // this issuer's group of the stream (may not exist in a tail round)
          const bool mine = g < G;
          if (mine) TC_HOT_WAIT(&a_ready[g], (a_it + (uint32_t)st) & 1);
          TC_T(9);
          const uint32_t a_lo = da_lo0 + (uint32_t)g * (TCA_A_BYTES >> 4);
          for (int tile = 0; tile < p.n_tiles; ++tile, ++x_it) {
            const int s = x_it % TCA_STAGES;
            TC_HOT_WAIT(&x_full[s], (x_it / TCA_STAGES) & 1);
            TC_T(8);
            if (mine) {
#pragma unroll
              for (int h = 0; h < TCA_HALVES; ++h) {  // rows [TCA_ACC_N h, TCA_ACC_N (h + 1)) of the tile -> one accumulator
                uint32_t& k = strm ? ks1 : ks0;       // accumulators this issuer has issued for the stream
                const int b = issuer * TCA_NBI + (int)(k % TCA_NBI);
                // TMEM buffer b must be drained by its previous users: this stream's accumulator k - NBI of this issuer and,
                // after a stream switch, the other stream's last one in this buffer.  uses(s, b) = #{j < k_s : j mod NBI = b'}.
                const uint32_t u_own = k / TCA_NBI, u_oth = ((strm ? ks0 : ks1) + (TCA_NBI - 1) - (k % TCA_NBI)) / TCA_NBI;
                if (u_own) TC_HOT_WAIT(&acc_empty[strm * TCA_NBUF + b], (u_own - 1) & 1);
                if (u_oth) TC_HOT_WAIT(&acc_empty[(strm ^ 1) * TCA_NBUF + b], (u_oth - 1) & 1);
                TC_T(10);
                tc_fence_after();
                // canonical K-major tile of 256 rows: row group r / 8 is 128 bytes further, so half h starts 16 groups in
                const uint32_t b_lo = db_lo0 + (uint32_t)s * (TCA_TILE_BYTES >> 4) + (uint32_t)h * (((TCA_ACC_N / 8) * 128) >> 4);
                const uint32_t d = tmem_base + (uint32_t)(b * TCA_ACC_N);
                // A chunks 0 1 | 2 3 | 0 1 (D_hi | D_lo | D_hi)  x  B chunks 0 1 | 0 1 | 2 3 (X_hi | X_hi | X_lo)
#ifndef AMCMC_TCA_K64
                umma_bf16_lean<false>(d, a_lo, da_hi, b_lo, db_hi, idesc);
#pragma unroll
                for (int ks = 1; ks < TCA_K / 16; ++ks)
                  umma_bf16_lean<true>(d, a_lo + ks * kAChunk, da_hi, b_lo + ks * kBChunk, db_hi, idesc);
#else
                umma_bf16_lean<false>(d, a_lo, da_hi, b_lo, db_hi, idesc);
                umma_bf16_lean<true>(d, a_lo + kAChunk, da_hi, b_lo + kBChunk, db_hi, idesc);
                if (!(ap.dbg & 2u)) {
                  umma_bf16_lean<true>(d, a_lo + 2 * kAChunk, da_hi, b_lo, db_hi, idesc);
                  umma_bf16_lean<true>(d, a_lo + 3 * kAChunk, da_hi, b_lo + kBChunk, db_hi, idesc);
                  umma_bf16_lean<true>(d, a_lo, da_hi, b_lo + 2 * kBChunk, db_hi, idesc);
                  umma_bf16_lean<true>(d, a_lo + kAChunk, da_hi, b_lo + 3 * kBChunk, db_hi, idesc);
                }
#endif
                umma_commit(&acc_full[strm * TCA_NBUF + b]);
                ++k;
              }
            }
            umma_commit(&x_empty[s]);  // (an issuer without a group in this stream arrives at once: the stage needs both)
            TC_T(11);
          }
        }
      }
      TC_ROUND_END
    } else if (warp == kDrawWarp) {
      TC_ROUND_BEGIN
      // ===== draw warp: the draws of every chain, one step ahead (arwmh.py:162-165,174) =====
      const int ht = lane;
      for (int64_t st = 0; st < p.n_steps; ++st) {
        const int64_t it = p.i0 + st;
        for (int g = 0; g < G; ++g) {
          mbar_wait(&v_empty[g], ((a_it + (uint32_t)st) & 1) ^ 1);
          for (int row = ht; row < TC_M; row += 32) {
            const int64_t c = (g0 + g * gs) * TC_M + row;
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
      TC_ROUND_END
    } else {
      // ===== epilogue / sampler threads =====
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSamplerRegs));
      TC_ROUND_BEGIN
      // Two streams: sampler warps 0-3 own groups 0 and 2, warps 4-7 groups 1 and 3.  The MMA warp serves the
      // streams alternately, so while one stream's likelihood runs on the tensor cores the other stream's warps
      // do their accept / factor pass / proposal (HBM- and issue-bound): the two phases overlap.
      const int q4 = warp & 3;
      const int half = warp >> 2;  // stream
      const int row = q4 * 32 + lane;
      const uint32_t t_lane = (uint32_t)(q4 * 32) << 16;
      const int n_mine = G > half ? (G - half + 1) / 2 : 0;
      // The two chains a thread owns (groups 2*half and 2*half+1) are served by ONE copy of the code: the loop
      // over them is rolled and the per-chain scalars are swapped at its end, so they stay in registers.
      float UcurA = 0.f, UcurB = 0.f, maccA = 0.f, maccB = 0.f, lamA = 0.f, lamB = 0.f, uaccA = 2.f, uaccB = 2.f;
      double UppA = 0.0, UppB = 0.0, i2vA = 0.0, i2vB = 0.0, rssA = 0.0, rssB = 0.0;
      int curA = 0, curB = 0;
      int64_t until_collect = p.collect_start + p.thinning;
      int64_t sidx = 0;
#pragma unroll 1
      for (int rep = 0; rep < 2; ++rep) {
        const int g = half + 2 * rep;
        if (g < G) {
          const int64_t c = (g0 + g * gs) * TC_M + row;
          const int64_t cc = c < p.C ? c : (p.C - 1);
          UcurA = p.pe[cc]; maccA = p.macc[cc]; lamA = ap.lam[cc];
        }
        { float tf; tf = UcurA; UcurA = UcurB; UcurB = tf; tf = maccA; maccA = maccB; maccB = tf; tf = lamA; lamA = lamB; lamB = tf; }
      }

#ifdef AMCMC_TC_TIMING
      const bool dbg = (blockIdx.x == 0 && lane == 0);
      long long tlast = clock64();
      unsigned long long* s_dbg_w = s_dbg + 16 + warp * 4;  // per sampler warp: GEMM phase, accept, v_full + pass, proposal
#endif
      // st = -1 builds the proposal of the first step (no accept, no update)
      for (int64_t st = -1; st < p.n_steps; ++st) {
        TC_W(3);
        const int64_t it = p.i0 + st;
        float mine0 = 0.f, mine1 = 0.f;
        if (st >= 0) {
          // ---- likelihood: sum_n m_n^2 of this stream's groups from the TMEM accumulators (256 columns each)
          for (int tile = 0; tile < p.n_tiles; ++tile) {
#pragma unroll
            for (int l = 0; l < 2; ++l) {  // l-th group of the stream <-> issuer l <-> TMEM buffers l * NBI ..
              if (l < n_mine) {
                float ss = 0.f;
#pragma unroll 1
                for (int h = 0; h < TCA_HALVES; ++h) {  // rolled: 128 live accumulator values at a time
                  uint32_t& cnt = l ? acc_it1 : acc_it;
                  const int b = l * TCA_NBI + (int)(cnt % TCA_NBI);
#ifdef AMCMC_TC_TIMING
                  const long long td0 = clock64();
#endif
                  TC_HOT_WAIT(&acc_full[half * TCA_NBUF + b], (cnt / TCA_NBI) & 1);
#ifdef AMCMC_TC_TIMING
                  const long long td1 = clock64();
#endif
                  tc_fence_after();
#pragma unroll 1
                  for (uint32_t ch = 0; ch < (uint32_t)TCA_ACC_N; ch += 128u)
                    ss += epilogue_sumsq_half(tmem_base + t_lane + (uint32_t)(b * TCA_ACC_N) + ch);
                  tc_fence_before();
                  // one arrival per warp: 128 per-thread arrivals on one mbarrier are 128 serialised shared-memory
                  // atomics per accumulator
                  __syncwarp();
                  if (lane == 0) mbar_arrive(&acc_empty[half * TCA_NBUF + b]);
#ifdef AMCMC_TC_TIMING
                  if (dbg) { s_dbg[48 + warp * 2] += (unsigned long long)(td1 - td0); s_dbg[49 + warp * 2] += (unsigned long long)(clock64() - td1); }
#endif
                  ++cnt;
                }
                if (l) mine1 += ss; else mine0 += ss;
              }
            }
          }
          TC_W(0);
          
        }
        bool collect_now = false;
        if (st >= 0) {
          collect_now = (--until_collect == 0);
          if (collect_now) until_collect = p.thinning;
        }
        const bool last = (st == p.n_steps - 1);
        const int64_t n = (it < ap.num_warmup) ? (it + 1) : (it + 1 - ap.num_warmup);
        const float nf = (float)n;
        const float gamma_n = (n == 1) ? 1.f : Num<float>::pow_neg(nf, ap.lr_decay);
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
          const int g = half + 2 * rep;
          if (g < G) {
            const int64_t c = (g0 + g * gs) * TC_M + row;
            const bool live = c < p.C;
            const int64_t cc = live ? c : (p.C - 1);
            float w[TC_D], acc[TC_D];
            float gamma = 0.f, el_old = 1.f, el_new;
            bool upd = false;
            if (st >= 0) {
              // ---- accept / reject (arwmh.py:170-178)
              const float mine = rep ? mine1 : mine0;
              // RSS is a sum of squares: assembled first and clamped at 0, so that the cancellation error of a far-away
              // proposal (the reference's start makes chains jump by hundreds, with e^{-2s} ~ 1e300) cannot turn into a
              // large NEGATIVE energy, which would be accepted and never left
              const double rss = rssA + (double)mine;
              float Up = (float)(UppA + i2vA * (rss > 0.0 ? rss : 0.0));
              if (Up != Up) Up = INFINITY;
              const float e = __expf(UcurA - Up);
              const float alpha = (e > 1.f) ? 1.f : e;
              const bool accd = uaccA < alpha;
              maccA = fmaf(alpha - maccA, Num<float>::rcp(nf), maccA);  // :185
              if (accd) { curA ^= 1; UcurA = Up; }
              const float* xs = curA ? p.xprop : p.z;
              if (live) {
                if (p.out_acc) p.out_acc[st * p.C + c] = (uint8_t)accd;
                if (collect_now) {
                  if (p.out_z) {
#pragma unroll
                    for (int k = 0; k < TC_D; ++k) p.out_z[(sidx * TC_D + k) * p.C + c] = xs[(int64_t)k * p.C + c];
                  }
                  if (p.out_pe) p.out_pe[sidx * p.C + c] = UcurA;
                }
              }
              // ---- adaptation (:188-193): mean, step size, rank-one factor update fused with the next proposal
              float dabs = 0.f;
              float mu[TC_D];
#pragma unroll
              for (int k = 0; k < TC_D; ++k) {  // all loads first: the stores below may alias them for the compiler
                mu[k] = ap.loc[(int64_t)k * p.C + cc];
                w[k] = xs[(int64_t)k * p.C + cc];
              }
#pragma unroll
              for (int k = 0; k < TC_D; ++k) {
                w[k] -= mu[k];
                if (live) ap.loc[(int64_t)k * p.C + c] = fmaf(gamma_n, w[k], mu[k]);
                dabs += fabsf(w[k]);
              }
              upd = live && (n != 1) && (dabs < Num<float>::kBig);  // padding lanes must never write a factor
              el_old = __expf(lamA);
              lamA = fmaf(gamma_n, alpha - ap.target, lamA);
              el_new = __expf(lamA);
            } else {
              el_new = __expf(lamA);
            }
            if (upd) {
              gamma = gamma_n;
            } else {
#pragma unroll
              for (int k = 0; k < TC_D; ++k) w[k] = 0.f;
            }
            const float* zn = sV + (size_t)g * 27 * TC_M + row;
            float* fcol = ap.ldl + (g0 + g * gs) * (TC_NE * TC_M) + row;
            TC_W(1);
            if (ap.dbg & 4u) {
              if (!last) mbar_wait(&v_full[g], (a_it + (uint32_t)(st + 1)) & 1);
#pragma unroll
              for (int k = 0; k < TC_D; ++k) acc[k] = 0.f;
            }
            if (ap.dbg & 4u) {
            } else if (last) {
              const float ss = tc_column_pass<true>(fcol, zn, false, upd, w, gamma, el_old, el_new, acc);
              if (live) ap.asc[c] = sqrtf(ss);  // :197
            } else {
              mbar_wait(&v_full[g], (a_it + (uint32_t)(st + 1)) & 1);
              
              tc_column_pass<false>(fcol, zn, true, upd, w, gamma, el_old, el_new, acc);
              TC_W(2);
            }
            if (!last) {
              // ---- proposal of step st+1 from the current position (arwmh.py:166-167), A' row, energy parts
              const float* xsrc = curA ? p.xprop : p.z;
              float* xdst = curA ? p.z : p.xprop;
              float xp[TC_D];
#pragma unroll
              for (int i = 0; i < TC_D; ++i) xp[i] = xsrc[(int64_t)i * p.C + cc];
#pragma unroll
              for (int i = 0; i < TC_D; ++i) {
                xp[i] += fmaf(el_new, acc[i], ap.eps * zn[i * TC_M]);
                if (live) xdst[(int64_t)i * p.C + c] = xp[i];
              }
              uaccA = zn[26 * TC_M];
              mbar_arrive(&v_empty[g]);
              tc_emit_proposal(xp, ap.cref + cc, p.C, ap.crss[cc], sA, g, row, &a_ready[g], (double)p.n_rows, p.cst, UppA, i2vA, rssA);
              TC_W(3);
            }
          }
          {  // swap the two owned chains
            float tf; double td; int ti;
            tf = UcurA; UcurA = UcurB; UcurB = tf;
            tf = maccA; maccA = maccB; maccB = tf;
            tf = lamA; lamA = lamB; lamB = tf;
            tf = uaccA; uaccA = uaccB; uaccB = tf;
            td = UppA; UppA = UppB; UppB = td;
            td = i2vA; i2vA = i2vB; i2vB = td;
            td = rssA; rssA = rssB; rssB = td;
            ti = curA; curA = curB; curB = ti;
          }
        }
        if (collect_now) ++sidx;
      }
      // ---- write back the per-chain scalars and the position if it ended in the shadow buffer
#pragma unroll 1
      for (int rep = 0; rep < 2; ++rep) {
        const int g = half + 2 * rep;
        if (g < G) {
          const int64_t c = (g0 + g * gs) * TC_M + row;
          if (c < p.C) {
            p.pe[c] = UcurA;
            p.macc[c] = maccA;
            ap.lam[c] = lamA;
            if (curA) {
#pragma unroll
              for (int k = 0; k < TC_D; ++k) p.z[(int64_t)k * p.C + c] = p.xprop[(int64_t)k * p.C + c];
            }
          }
        }
        float tf; int ti;
        tf = UcurA; UcurA = UcurB; UcurB = tf;
        tf = maccA; maccA = maccB; maccB = tf;
        tf = lamA; lamA = lamB; lamB = tf;
        ti = curA; curA = curB; curB = ti;
      }
      TC_ROUND_END
    }
  }
#undef TC_ROUND_BEGIN
#undef TC_ROUND_END
#undef TC_SERVICE_REGS
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
#ifdef AMCMC_TC_TIMING
  if (blockIdx.x == 0 && tid < 64) atomicAdd(&tc_dbg[tid], s_dbg[tid]);
#endif
}

// from diamonds_tc.cu
int diamonds_tc_prepare(const amcmc_model* m, int64_t n_chains, TcParams* p, const amcmc_state* st, const amcmc_run_args* a);
void diamonds_tc_launch_ref(const amcmc_model* m, const float* loc, const float* scale, const float* lam, double eps, cudaStream_t s);

// Adaptive run on the tensor cores: the full ARWMH.sample for every chain of *st.
int run_diamonds_tc_adapt(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  DiamondsTcExtra* ex = (DiamondsTcExtra*)m->extra;
  if (!ex) { set_error("diamonds tensor-core path needs K = 25 predictors and fp32"); return AMCMC_ERR_UNSUPPORTED; }
  TcAdaptParams ap;
  tc_run_begin(ex, s);
  int rc = diamonds_tc_prepare(m, st->n_chains, &ap.p, st, a);
  if (rc) return rc;
  ap.loc = (float*)st->loc;
  ap.lam = (float*)st->log_step_size;
  ap.asc = (float*)st->as_change;
  ap.num_warmup = a->num_warmup;
  ap.lr_decay = (float)a->lr_decay;
  ap.target = (float)a->target_accept_prob;
  ap.eps = (float)a->eps;
  ap.x_bytes = TCA_TILE_BYTES;
  if (const char* e = getenv("AMCMC_TC_EXPERIMENT_X_BYTES")) {  // traffic experiment only (results are then wrong)
    const long v = atol(e);
    if (v >= 16 && v <= TCA_TILE_BYTES && v % 16 == 0) ap.x_bytes = (uint32_t)v;
  }
  ap.dbg = 0;
  if (const char* e = getenv("AMCMC_TC_EXPERIMENT_FLAGS")) ap.dbg = (uint32_t)atol(e);
  const int64_t C = st->n_chains;
  if (ex->ldl_groups < ap.p.n_groups) {
    if (ex->ldl) cudaFree(ex->ldl);
    ex->ldl = nullptr;
    ex->ldl_groups = 0;
    if ((rc = check_cuda(cudaMalloc(&ex->ldl, (size_t)ap.p.n_groups * TC_NE * TC_M * sizeof(float)), "cudaMalloc(ldl)"))) return rc;
    ex->ldl_groups = ap.p.n_groups;
  }
  ap.ldl = ex->ldl;
  if (ex->cref_cap < C) {
    if (ex->cref) cudaFree(ex->cref);
    if (ex->crss) cudaFree(ex->crss);
    ex->cref = nullptr; ex->crss = nullptr; ex->cref_cap = 0;
    if ((rc = check_cuda(cudaMalloc(&ex->cref, (size_t)C * 2 * TC_KC * sizeof(float)), "cudaMalloc(cref)"))) return rc;
    if ((rc = check_cuda(cudaMalloc(&ex->crss, (size_t)C * sizeof(double)), "cudaMalloc(crss)"))) return rc;
    ex->cref_cap = C;
  }
  ap.cref = ex->cref;
#ifndef AMCMC_TCA_K64
  ap.Xcanon64 = ex->Xcanon;   // the K = 80 tiles of the shared-state kernel
#else
  ap.Xcanon64 = ex->Xcanon64;
#endif
  ap.crss = ex->crss;
  tc_chol_to_ldl_kernel<<<ap.p.n_groups, TC_M, 0, s>>>((const float*)st->scale, ap.ldl, C);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (const char* cap = getenv("AMCMC_TC_MAX_CTAS")) {  // test hook: several groups / rounds per CTA at small chain counts
    const int v = atoi(cap);
    if (v > 0 && v < sms) sms = v;
  }
  const int grid = ap.p.n_groups < sms ? ap.p.n_groups : sms;
  if ((rc = check_cuda(cudaFuncSetAttribute(diamonds_tc_adapt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcaSmem::BYTES), "cudaFuncSetAttribute"))) return rc;
  if ((rc = check_cuda(cudaFuncSetAttribute(diamonds_tc_adapt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcaSmem::BYTES), "cudaFuncSetAttribute"))) return rc;
  // The centred GEMM is accurate while a chain is close to its reference point (the cancellation in
  // RSS_ref - 2 d.g + sum m^2 grows with |d|^2), so the run is cut into segments of at most kSegment steps and every
  // chain's reference point is moved to its current position before each of them.  Per-chain state stays in the
  // LDL^T blocks across the segments.
  // L2 residency: at 65,536 chains the 92 MB of factors plus ~35 MB of positions / means / reference points exceed
  // the 126 MB L2 by a little, and a cyclic sweep through slightly-too-much data gets no hits from an LRU-like policy
  // (ncu: the whole factor is re-streamed from HBM every step).  The factors of the first two groups of every CTA
  // (a contiguous prefix of the buffer, see the kernel) are pinned with an access-policy window; the rest streams.
  // (when a CTA serves more than TC_GR groups it does so in rounds; every round is its own launch, with the window
  // moved to the first two groups of that round -- group ids of local index l are the contiguous range [l grid, (l+1) grid))
  int max_persist = 0, max_window = 0;
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  const size_t group_bytes = (size_t)TC_NE * TC_M * sizeof(float);
  const size_t total_bytes = (size_t)ap.p.n_groups * group_bytes;
  double per_cta = 2.0;  // pinned groups per CTA (tunable for experiments: AMCMC_TC_PIN_GROUPS)
  if (const char* e = getenv("AMCMC_TC_PIN_GROUPS")) per_cta = atof(e);
  const bool want_pin = total_bytes > ((size_t)48 << 20) && !getenv("AMCMC_TC_NO_L2_PIN");  // small batches fit in L2 anyway
  bool pinned = false;
  int pinned_round = -1;
  auto pin_round = [&](int rnd) {
    if (!want_pin || rnd == pinned_round) return;  // (a single-round run sets the window once for all its segments)
    pinned_round = rnd;
    const size_t first = (size_t)rnd * TC_GR * grid;                 // first group of the round
    if (first >= (size_t)ap.p.n_groups) return;
    size_t want = (size_t)(per_cta * grid) * group_bytes;
    if (first * group_bytes + want > total_bytes) want = total_bytes - first * group_bytes;
    if (want > (size_t)max_persist) want = (size_t)max_persist;
    if (want > (size_t)max_window) want = (size_t)max_window;
    if (want == 0) return;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = (char*)ap.ldl + first * group_bytes;
    attr.accessPolicyWindow.num_bytes = want;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (!pinned) {
      if (!ex->l2_dirty) {  // remember the caller's limit; tc_release_l2 restores it once this run has finished
        size_t prev = 0;
        cudaDeviceGetLimit(&prev, cudaLimitPersistingL2CacheSize);
        ex->l2_prev_limit = prev;
      }
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); return; }
      ex->l2_dirty = 1;
    }
    if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) pinned = true;
    else cudaGetLastError();  // pinning is an optimisation: run without it
  };
  const int rounds_max = ((ap.p.n_groups + grid - 1) / grid + TC_GR - 1) / TC_GR;
  const TcParams full = ap.p;
  int64_t seg = kSegment;
  if (const char* e = getenv("AMCMC_TC_SEGMENT")) {  // test hook: exercise the segment bookkeeping in short runs
    const long v = atol(e);
    if (v > 0) seg = v;
  }
  for (int64_t s0 = 0; s0 < full.n_steps; s0 += seg) {
    TcParams& p = ap.p;
    p = full;
    p.i0 = full.i0 + s0;
    p.n_steps = full.n_steps - s0 < seg ? full.n_steps - s0 : seg;
    // sample bookkeeping of the segment: steps until the next kept sample, index of that sample
    const int64_t r = s0 - full.collect_start;
    const int64_t kept = r <= 0 ? 0 : r / full.thinning;
    p.collect_start = r <= 0 ? -r : -(r % full.thinning);
    if (full.out_z) p.out_z = full.out_z + kept * TC_D * C;
    if (full.out_pe) p.out_pe = full.out_pe + kept * C;
    if (full.out_acc) p.out_acc = full.out_acc + s0 * C;
    if (full.normals) p.normals = full.normals + s0 * TC_D * C;
    if (full.uniforms) p.uniforms = full.uniforms + s0 * C;
    tc_chain_ref_kernel<<<(unsigned)((C + 127) / 128), 128, 0, s>>>(ex->gram, (const float*)st->z, C, ex->cref, ex->crss);
    for (int rnd = 0; rnd < rounds_max; ++rnd) {
      pin_round(rnd);
      ap.rnd_begin = rnd;
      ap.rnd_end = rnd + 1;
      if (a->rng_mode == AMCMC_RNG_EXTERNAL) diamonds_tc_adapt_kernel<true><<<grid, TC_THREADS, TcaSmem::BYTES, s>>>(ap);
      else diamonds_tc_adapt_kernel<false><<<grid, TC_THREADS, TcaSmem::BYTES, s>>>(ap);
      if ((rc = check_cuda(cudaGetLastError(), "diamonds_tc_adapt_kernel launch"))) return rc;
    }
  }
  tc_ldl_to_chol_kernel<<<ap.p.n_groups, TC_M, 0, s>>>(ap.ldl, (float*)st->scale, C);
  rc = check_cuda(cudaGetLastError(), "tc_ldl_to_chol_kernel launch");
  if (pinned) {  // drop the window for whatever runs next on this stream
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
  }
  tc_run_end(ex, s);  // the persisting lines themselves are released by tc_release_l2 when this event has completed
  return rc;
}

}  // namespace amcmc

#ifdef AMCMC_TC_TIMING
extern "C" void amcmc_debug_tc_timing(unsigned long long* out, int reset) {
  if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(amcmc::tc_dbg, z, sizeof(z)); }
  else cudaMemcpyFromSymbol(out, amcmc::tc_dbg, sizeof(unsigned long long) * 64);
}
#endif
