// launch_block.cuh -- host-side launch helpers for the block-per-chain kernels.
#pragma once
#include <cstdlib>
#include "arwmh_block.cuh"
#include "asss_block.cuh"
#include "launch_small.cuh"

namespace amcmc {

template <class K> inline int ensure_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return AMCMC_OK;
  return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                    "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
}

template <class BM, typename R>
int launch_block_run(const BM& m, int d, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  if (d > 32) {
    set_error("block kernel: per-chain adaptation is compiled for d <= 32 (got %d)", d);
    return AMCMC_ERR_UNSUPPORTED;
  }
  const StateView<R> sv = make_state_view<R>(st);
  const RunView<R> rv = make_run_view<R>(st, a);
  const size_t smem = BlockSmem<R>::bytes(d);
  const unsigned grid = (unsigned)st->n_chains;
  const bool ext = a->rng_mode == AMCMC_RNG_EXTERNAL;
  int rc = AMCMC_OK;
  // Few chains (at most one per SM): one 512-thread CTA per chain.  The likelihood phase is bound by the loads a CTA
  // keeps in flight; measured per step at 64 diamonds chains: 256 threads 14.9 us, 512 threads 10.4 us, 1024 threads
  // 12.5 us (the 64-register cap slows the warp-0 sweep).  With two or more chains per SM the 256-thread CTAs win
  // (200 chains: 15.8 us vs 20.3 us).  Staging half of the design matrix in the shared memory a wide CTA leaves unused
  // was measured as well and changed nothing.
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool wide = st->n_chains <= (int64_t)kWideMaxChainsPerSm * sms && !getenv("AMCMC_BLOCK_NARROW");
  if (a->kernel_kind == AMCMC_KERNEL_ASSS) {
#define AMCMC_LA(EX, AD)                                                              \
  do {                                                                                \
    if (wide) {                                                                       \
      auto k = asss_block_kernel<BM, R, EX, AD, kBlockThreadsWide>;                   \
      if ((rc = ensure_smem(k, smem))) return rc;                                     \
      k<<<grid, kBlockThreadsWide, smem, s>>>(m, sv, rv, d);                          \
    } else {                                                                          \
      auto k = asss_block_kernel<BM, R, EX, AD, kBlockThreads>;                       \
      if ((rc = ensure_smem(k, smem))) return rc;                                     \
      k<<<grid, kBlockThreads, smem, s>>>(m, sv, rv, d);                              \
    }                                                                                 \
  } while (0)
    if (a->adapt) { if (ext) AMCMC_LA(true, true); else AMCMC_LA(false, true); }
    else          { if (ext) AMCMC_LA(true, false); else AMCMC_LA(false, false); }  // frozen: ASSS.sample_Pnx
#undef AMCMC_LA
    return check_cuda(cudaGetLastError(), "asss_block_kernel launch");
  }
#define AMCMC_LB(AD, EX)                                                              \
  do {                                                                                \
    if (wide) {                                                                       \
      auto k = arwmh_block_kernel<BM, R, AD, EX, kBlockThreadsWide>;                  \
      if ((rc = ensure_smem(k, smem))) return rc;                                     \
      k<<<grid, kBlockThreadsWide, smem, s>>>(m, sv, rv, d);                          \
    } else {                                                                          \
      auto k = arwmh_block_kernel<BM, R, AD, EX, kBlockThreads>;                      \
      if ((rc = ensure_smem(k, smem))) return rc;                                     \
      k<<<grid, kBlockThreads, smem, s>>>(m, sv, rv, d);                              \
    }                                                                                 \
  } while (0)
  if (a->adapt) { if (ext) AMCMC_LB(true, true); else AMCMC_LB(true, false); }
  else          { if (ext) AMCMC_LB(false, true); else AMCMC_LB(false, false); }
#undef AMCMC_LB
  return check_cuda(cudaGetLastError(), "arwmh_block_kernel launch");
}

template <class BM, typename R>
int launch_block_init(const BM& m, int d, const amcmc_state* st, uint64_t seed, int64_t chain_offset, double radius,
                      int use_given_z, cudaStream_t s) {
  const StateView<R> sv = make_state_view<R>(st);
  const size_t smem = sizeof(R) * (((d + 3) & ~3) + 40);
  arwmh_block_init_kernel<BM, R><<<(unsigned)st->n_chains, kBlockThreads, smem, s>>>(m, sv, d, seed, chain_offset,
                                                                                    (R)radius, use_given_z);
  return check_cuda(cudaGetLastError(), "arwmh_block_init_kernel launch");
}

template <class BM, typename R>
int launch_block_potential(const BM& m, int d, int64_t n, const void* q, void* out, cudaStream_t s) {
  const size_t smem = sizeof(R) * (((d + 3) & ~3) + 40);
  potential_block_kernel<BM, R><<<(unsigned)n, kBlockThreads, smem, s>>>(m, d, n, (const R*)q, (R*)out);
  return check_cuda(cudaGetLastError(), "potential_block_kernel launch");
}

}  // namespace amcmc
