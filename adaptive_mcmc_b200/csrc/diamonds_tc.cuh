// diamonds_tc.cuh -- constants, parameter blocks and device helpers shared by the diamonds tensor-core kernels
// (diamonds_tc.cu: shared adaptation state; diamonds_tc_adapt.cu: per-chain adaptation).
#pragma once
#include <cmath>
#include <cstring>
#include "common.cuh"
#include "internal.h"
#include "tc_common.cuh"

namespace amcmc {

using namespace tc;

constexpr int TC_D = 26;          // model dimension
constexpr int TC_KC = 25;         // [1 | Xc] columns
constexpr int TC_KP = 80;         // concatenated split-precision K
constexpr int TC_TILE_N = 256;    // data rows per tile (UMMA N)
constexpr int TC_M = 128;         // chains per group (UMMA M)
constexpr int TC_GR = 4;          // groups per round
constexpr int TC_EPI_WARPS = 8;     // 2 per TMEM lane quarter: each drains half the columns of every tile
constexpr int TC_HELP_WARPS = 2;    // precompute the state-independent half of the next proposal
constexpr int TC_THREADS = 32 * (2 + TC_EPI_WARPS + TC_HELP_WARPS);  // 384
constexpr int TC_TILE_BYTES = TC_TILE_N * TC_KP * 2;  // 40960
constexpr int TC_A_BYTES = TC_M * TC_KP * 2;          // 20480
constexpr int TC_NP = TC_D * (TC_D + 1) / 2;          // 351

// layout of the per-window reference block (floats)
constexpr int REF_Q = 0;     // q_ref[26]
constexpr int REF_G2 = 32;   // 2*g[25]
constexpr int REF_RSS = 60;  // RSS_ref as a float64 (two float slots, 8-byte aligned)
constexpr int REF_S = 64;    // S = e^lam L + eps I, dense lower rows padded to 28 floats (zeros above the diagonal)
constexpr int REF_SLD = 28;  // row stride of S: 16-byte vector loads, no triangular guards
constexpr int REF_FLOATS = 64 + TC_D * REF_SLD;

struct DiamondsTcExtra {
  uint16_t* Xcanon;  // [n_tiles][TC_TILE_BYTES/2] bf16, canonical UMMA tile order
  uint16_t* Xcanon64;  // [n_tiles][256 x 64] bf16: [X_hi (25 + 7 zeros) | X_lo (25 + 7 zeros)], for the per-chain-adaptive kernel
  int n_tiles;
  double* gram;      // G[25*25] = X1^T X1, h[25] = X1^T Y, yy   (fp64, device)
  float* ref;        // [REF_FLOATS] device
  float* xprop;      // [26][cap] proposal parking buffer
  int64_t xprop_cap;
  // per-chain-adaptive path: reference point = mean position of the batch
  double* mean_acc;  // [32]
  float* qmean;      // [26]
  float* ident;      // [351] dummy packed scale for the reference-block kernel
  float* zero;       // [1]
  float* ldl;        // [ldl_groups][351][128] per-chain LDL^T factors of the adaptive path (see diamonds_tc_adapt.cu)
  int64_t ldl_groups;
  float* cref;       // [50][cref_cap] per-chain GEMM reference: rows 0-24 q_ref, rows 25-49 2 g
  double* crss;      // [cref_cap] RSS at the reference point
  int64_t cref_cap;
  // The scratch above belongs to the model handle, so runs on one handle are serialised: a run on another stream waits
  // for `done_ev` of the previous one (tc_run_begin / tc_run_end).  Host threads must not share a handle concurrently.
  cudaEvent_t done_ev;
  cudaStream_t last_stream;
  int have_done_ev;
  // persisting-L2 window of the adaptive path: released (cudaCtxResetPersistingL2Cache + previous limit restored) once
  // the run that set it has finished -- checked at the next call on the handle and at amcmc_model_destroy
  int l2_dirty;
  size_t l2_prev_limit;
};

void tc_run_begin(DiamondsTcExtra* ex, cudaStream_t s);  // serialise against the previous run; release a finished L2 window
void tc_run_end(DiamondsTcExtra* ex, cudaStream_t s);
void tc_release_l2(DiamondsTcExtra* ex, bool wait);

struct TcParams {
  int64_t C;
  int n_groups;
  const uint16_t* Xcanon;
  int n_tiles;
  const float* ref;
  float* z;      // [26][C]
  float* pe;     // [C]
  float* macc;   // [C] mean accept probability over this launch
  float* xprop;  // [26][C]
  int64_t i0, n_steps, thinning, collect_start;
  uint64_t seed;
  int64_t chain_offset;
  const float* normals;   // EXTERNAL [T][26][C]
  const float* uniforms;  // [T][C]
  float* out_z;
  float* out_pe;
  uint8_t* out_acc;
  float n_rows;
  double cst;
};

__device__ __forceinline__ uint16_t f2bf(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf2f(uint16_t b) { return __bfloat162float(__ushort_as_bfloat16(b)); }


struct TcSmem {
  static constexpr int OFF_X = 0;                              // 2 stages
  static constexpr int OFF_A = 2 * TC_TILE_BYTES;              // TC_GR groups
  static constexpr int OFF_REF = OFF_A + TC_GR * TC_A_BYTES;   // REF_FLOATS floats
  static constexpr int OFF_BAR = OFF_REF + REF_FLOATS * 4;     // barriers
  static constexpr int N_BAR = 2 + 2 + 2 + 2 + 3 * TC_GR;      // x_full, x_empty, acc_full, acc_empty, a_ready, v_full, v_empty
  static constexpr int OFF_TMEM = OFF_BAR + N_BAR * 8;
  static constexpr int OFF_EXCH = OFF_TMEM + 16;               // float [TC_GR][TC_M]: partner's half of sum m^2
  static constexpr int OFF_V = OFF_EXCH + TC_GR * TC_M * 4;    // float [TC_GR][27][TC_M]
  static constexpr int BYTES = OFF_V + TC_GR * 27 * TC_M * 4;
};

// packed fp32x2 FMA on (lo, hi) register pairs: acc += v*v for two accumulator columns at once (FFMA2)
__device__ __forceinline__ void sq_acc2(float& a_lo, float& a_hi, float v_lo, float v_hi) {
  asm("{\n\t.reg .b64 rv, ra;\n\t"
      "mov.b64 rv, {%2, %3};\n\tmov.b64 ra, {%0, %1};\n\t"
      "fma.rn.f32x2 ra, rv, rv, ra;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a_lo), "+f"(a_hi)
      : "f"(v_lo), "f"(v_hi));
}

// sum of squares of this thread's TMEM lane over 128 accumulator columns: 4 loads in flight, one wait
__device__ __forceinline__ float epilogue_sumsq_half(uint32_t taddr) {
  float v0[32], v1[32], v2[32], v3[32];
  tmem_ld_32x32(taddr, v0);
  tmem_ld_32x32(taddr + 32u, v1);
  tmem_ld_32x32(taddr + 64u, v2);
  tmem_ld_32x32(taddr + 96u, v3);
  tmem_ld_wait();
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    sq_acc2(a0, a1, v0[i], v0[i + 1]);
    sq_acc2(a2, a3, v1[i], v1[i + 1]);
    sq_acc2(a4, a5, v2[i], v2[i + 1]);
    sq_acc2(a6, a7, v3[i], v3[i + 1]);
  }
  return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}


}  // namespace amcmc
