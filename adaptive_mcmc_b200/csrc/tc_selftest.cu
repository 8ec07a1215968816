// tc_selftest.cu -- hardware self-test of the tcgen05 building blocks used by diamonds_tc.cu:
// canonical no-swizzle K-major smem tiles, UMMA descriptors, TMEM alloc/ld, mbarrier + TMA bulk copy.
// D[128 x 256] = A[128 x K] * B[256 x K]^T (bf16 in, fp32 accumulate), one CTA.
#include "internal.h"
#include "tc_common.cuh"

namespace amcmc {

using namespace tc;

// rearrange a row-major [rows x K] bf16 matrix into the canonical tile order (global -> global)
__global__ void canon_tile_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int rows, int K) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * K) return;
  const int r = idx / K, k = idx % K;
  dst[canon_off(r, k, rows) / 2] = src[idx];
}

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, const uint16_t* __restrict__ Bcanon,
                     int K, float* __restrict__ D, int use_tma, int swap) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int M = 128, N = 256;
  uint16_t* sA = reinterpret_cast<uint16_t*>(smem);                  // M*K*2 bytes
  uint16_t* sB = reinterpret_cast<uint16_t*>(smem + M * 96 * 2);     // N*K*2 bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M * 96 * 2 + N * 96 * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bars[0], 1);  // MMA done
    mbar_init(&bars[1], 1);  // TMA landed
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  // stage A (and B unless TMA) into the canonical layout with ordinary stores
  for (int idx = tid; idx < M * K; idx += 128) sA[canon_off(idx / K, idx % K, M) / 2] = A[idx];
  if (!use_tma)
    for (int idx = tid; idx < N * K; idx += 128) sB[canon_off(idx / K, idx % K, N) / 2] = B[idx];
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (use_tma && tid == 0) {
    mbar_arrive_expect_tx(&bars[1], (uint32_t)(N * K * 2));
    tma_load_1d(sB, Bcanon, (uint32_t)(N * K * 2), &bars[1]);
  }
  if (tid == 0) {
    if (use_tma) mbar_wait(&bars[1], 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16_f32(M, N);
    const uint32_t a_kstride = (M / 8) * 128, b_kstride = (N / 8) * 128;  // bytes between K chunks
    for (int s = 0; s < K / 16; ++s) {
      const uint32_t a_addr = smem_u32(sA) + 2 * s * a_kstride;
      const uint32_t b_addr = smem_u32(sB) + 2 * s * b_kstride;
      const uint64_t da = swap ? make_smem_desc(a_addr, 128, a_kstride) : make_smem_desc(a_addr, a_kstride, 128);
      const uint64_t db = swap ? make_smem_desc(b_addr, 128, b_kstride) : make_smem_desc(b_addr, b_kstride, 128);
      umma_bf16(tmem_base, da, db, idesc, s > 0);
    }
    umma_commit(&bars[0]);
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();
  // epilogue: warp w reads TMEM lanes 32w..32w+31
  const int row = tid;
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    // exercise the packed FFMA2 path too: v*1 + 0
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float2 r = ffma2(make_float2(v[i], v[i + 1]), make_float2(1.f, 1.f), make_float2(0.f, 0.f));
      D[row * N + c0 + i] = r.x;
      D[row * N + c0 + i + 1] = r.y;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace amcmc

extern "C" int amcmc_selftest_umma(const void* a_bf16, const void* b_bf16, int K, void* scratch, float* d_out,
                                   int use_tma, int swap_lbo_sbo, void* stream) {
  using namespace amcmc;
  if (!a_bf16 || !b_bf16 || !d_out || K < 16 || K > 96 || (K % 16) != 0 || (use_tma && !scratch)) {
    set_error("amcmc_selftest_umma: bad arguments (K must be a multiple of 16 in [16, 96])");
    return AMCMC_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = 128 * 96 * 2 + 256 * 96 * 2 + 64;
  int rc = check_cuda(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "cudaFuncSetAttribute");
  if (rc) return rc;
  if (use_tma)
    canon_tile_kernel<<<(256 * K + 255) / 256, 256, 0, s>>>((const uint16_t*)b_bf16, (uint16_t*)scratch, 256, K);
  umma_selftest_kernel<<<1, 128, smem, s>>>((const uint16_t*)a_bf16, (const uint16_t*)b_bf16, (const uint16_t*)scratch, K,
                                            d_out, use_tma, swap_lbo_sbo);
  return check_cuda(cudaGetLastError(), "umma_selftest_kernel launch");
}
