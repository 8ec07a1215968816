// umma_bench.cu -- micro-benchmark (probe, not product code): how many SM cycles does one tcgen05.mma.kind::f16 of the shape
// the diamonds kernels issue (M = 128, K = 16, N = 256 or 128, bf16 operands in the canonical NO-SWIZZLE K-major layout) take
// when issued back to back, and how long does the issuing thread spend issuing?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I adaptive_mcmc_b200/csrc -o scripts/probes/umma_bench.bin scripts/probes/umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace amcmc::tc;

__global__ void __launch_bounds__(128, 1) bench(int N, int n_acc, int mma_per_acc, int commit_each, int mode, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                 // 128 x 64 bf16 = 16 KB
  unsigned char* sB = smem + 16384;         // 256 x 64 bf16 = 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16_f32(128, N);
    const uint32_t a_ks = (128 / 8) * 128, b_ks = (256 / 8) * 128;
    const int nbuf = 512 / N;
    const long long t0 = clock64();
    uint32_t phase = 0;
    if (mode == 5) {  // descriptors hoisted, six MMAs unrolled with compile-time accumulate flags: the leanest issue loop
      uint64_t da[4], db[4];
      for (int c = 0; c < 4; ++c) {
        da[c] = make_smem_desc(smem_u32(sA) + 2 * c * a_ks, a_ks, 128);
        db[c] = make_smem_desc(smem_u32(sB) + 2 * c * b_ks, b_ks, 128);
      }
      for (int a = 0; a < n_acc; ++a) {
        const uint32_t d = tmem + (uint32_t)((a & (nbuf - 1)) * N);
        umma_bf16(d, da[0], db[0], idesc, false);
        umma_bf16(d, da[1], db[1], idesc, true);
        umma_bf16(d, da[2], db[0], idesc, true);
        umma_bf16(d, da[3], db[1], idesc, true);
        umma_bf16(d, da[0], db[2], idesc, true);
        umma_bf16(d, da[1], db[3], idesc, true);
        if (commit_each) umma_commit(&bar[1]);  // per-accumulator commit (as acc_full in the kernel), never waited on here
      }
    } else
    for (int a = 0; a < n_acc; ++a) {
      // mode 0: alternate buffers, first MMA overwrites.  1: same buffer always.  2: alternate buffers, always accumulate.
      // 3: two accumulators interleaved MMA by MMA (independent neighbours).  4: same buffer, always accumulate.
      const int b = (mode == 1 || mode == 4) ? 0 : a % nbuf;
      for (int ks = 0; ks < mma_per_acc; ++ks) {
        const int ia = (ks % 6) < 4 ? (ks % 6) : (ks % 6) - 4, ib = (ks % 6) < 2 ? (ks % 6) : (ks % 6) - 2;
        const uint64_t da = make_smem_desc(smem_u32(sA) + 2 * ia * a_ks, a_ks, 128);
        const uint64_t db = make_smem_desc(smem_u32(sB) + 2 * ib * b_ks, b_ks, 128);
        const bool accum = (mode == 2 || mode == 4) ? true : ks > 0;
        umma_bf16(tmem + (uint32_t)(b * N), da, db, idesc, accum);
        if (mode == 3) umma_bf16(tmem + (uint32_t)(((b + 1) % nbuf) * N), da, db, idesc, accum);
      }
      if (commit_each == 1) umma_commit(&bar[1]);
      if (commit_each == 2) { umma_commit(&bar[0]); mbar_wait(&bar[0], phase & 1); ++phase; }
    }
    const long long t1 = clock64();
    if (commit_each != 2) {
      umma_commit(&bar[0]);
      mbar_wait(&bar[0], phase & 1);
    }
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t2 - t0); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 64);
  const int n_acc = 2000;
  for (int N : {256, 128})
    for (int mode : {0, 5})
      for (int ce : {0, 1}) {
        const int mpa = 6;
        bench<<<148, 128, 16384 + 32768 + 64>>>(N, n_acc, mpa, ce, mode, d);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("N %3d mode %d commit-per-acc %d: issue %.1f cyc/MMA, total %.1f cyc/MMA, %.0f cyc/accumulator  (%s)\n", N, mode, ce,
               (double)h[0] / (n_acc * mpa), (double)h[1] / (n_acc * mpa), (double)h[1] / n_acc, cudaGetErrorString(e));
      }
  return 0;
}
