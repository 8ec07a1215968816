// launch_block.cuh -- host-side launch helpers for the block-per-chain kernels.
#pragma once
#include <cstdlib>
#include <type_traits>
#include "arwmh_block.cuh"
#include "asss_block.cuh"
#include "launch_small.cuh"

namespace amcmc {

// models whose likelihood is a sum over data rows that two CTAs can share declare `static constexpr bool kRowSplit = true`
// and provide rss_rows<NT>(q, red, r_begin, r_end) / finish(q, rss)
template <class BM, class = void> struct has_row_split : std::false_type {};
template <class BM> struct has_row_split<BM, std::void_t<decltype(BM::kRowSplit)>> : std::bool_constant<BM::kRowSplit> {};

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
  // Few-chain diamonds (at most one chain per TWO SMs): a thread-block cluster per chain splits the data rows of the
  // likelihood (arwmh_block.cuh, CL = 2, 4, 8).  AMCMC_BLOCK_CLUSTER=0/2/4/8 overrides.
  if constexpr (has_row_split<BM>::value) {
    const char* cl_e = getenv("AMCMC_BLOCK_CLUSTER");  // read per launch: tests switch it inside one process
    const int cl_env = cl_e ? atoi(cl_e) : -1;
    // CTAs per chain: the largest cluster size whose clusters are all co-resident (a cluster lives inside one GPC, so fewer
    // clusters of 8 fit than SMs / 8: measured 18 chains x 8 CTAs = two waves, 14.1 us per step instead of 7.1);
    // cudaOccupancyMaxActiveClusters answers per cluster size, cached per device.  The environment variable gives the
    // cluster size directly (0 or 1: none).
    static int max_clusters[64][3] = {};  // [device][CL = 2, 4, 8], 0 = not asked yet, -1 = unavailable
    int cl = 1;
    if (a->n_steps > 0 && dev >= 0 && dev < 64) {
      const int sizes[3] = {2, 4, 8};
      for (int k = 2; k >= 0 && cl == 1; --k) {
        if (max_clusters[dev][k] == 0) {
          cudaLaunchConfig_t q = {};
          q.gridDim = dim3((unsigned)(sizes[k] * sms));
          q.blockDim = dim3(kBlockThreadsWide);
          q.dynamicSmemBytes = smem;
          cudaLaunchAttribute qa[1];
          qa[0].id = cudaLaunchAttributeClusterDimension;
          qa[0].val.clusterDim.x = (unsigned)sizes[k];
          qa[0].val.clusterDim.y = 1;
          qa[0].val.clusterDim.z = 1;
          q.attrs = qa;
          q.numAttrs = 1;
          int nmax = 0;
          cudaError_t e = cudaSuccess;
          if (ensure_smem(arwmh_block_kernel<BM, R, true, false, kBlockThreadsWide, 2>, smem) != AMCMC_OK) e = cudaErrorUnknown;
          if (e == cudaSuccess) {
            if (sizes[k] == 8) e = cudaOccupancyMaxActiveClusters(&nmax, arwmh_block_kernel<BM, R, true, false, kBlockThreadsWide, 8>, &q);
            else if (sizes[k] == 4) e = cudaOccupancyMaxActiveClusters(&nmax, arwmh_block_kernel<BM, R, true, false, kBlockThreadsWide, 4>, &q);
            else e = cudaOccupancyMaxActiveClusters(&nmax, arwmh_block_kernel<BM, R, true, false, kBlockThreadsWide, 2>, &q);
          }
          if (e != cudaSuccess) { cudaGetLastError(); nmax = -1; }
          max_clusters[dev][k] = nmax > 0 ? nmax : -1;
        }
        if (max_clusters[dev][k] >= st->n_chains) cl = sizes[k];
      }
    }
    if (cl_env >= 0) cl = (cl_env == 2 || cl_env == 4 || cl_env == 8) ? cl_env : 1;
    if (cl > 1 && a->n_steps > 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)cl * grid);
      cfg.blockDim = dim3(kBlockThreadsWide);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)cl;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
#define AMCMC_LC1(KERN, T1, T2, CLN)                                                   \
  do {                                                                                \
    auto k = KERN<BM, R, T1, T2, kBlockThreadsWide, CLN>;                             \
    if ((rc = ensure_smem(k, smem))) return rc;                                       \
    rc = check_cuda(cudaLaunchKernelEx(&cfg, k, m, sv, rv, d), #KERN " (cluster) launch"); \
  } while (0)
#define AMCMC_LC(KERN, T1, T2)                                                        \
  do {                                                                                \
    if (cl == 8) AMCMC_LC1(KERN, T1, T2, 8); else if (cl == 4) AMCMC_LC1(KERN, T1, T2, 4); else AMCMC_LC1(KERN, T1, T2, 2); \
  } while (0)
      if (a->kernel_kind == AMCMC_KERNEL_ASSS) {  // template order <EXTERNAL, ADAPT>
        if (a->adapt) { if (ext) AMCMC_LC(asss_block_kernel, true, true); else AMCMC_LC(asss_block_kernel, false, true); }
        else          { if (ext) AMCMC_LC(asss_block_kernel, true, false); else AMCMC_LC(asss_block_kernel, false, false); }
      } else {                                    // template order <ADAPT, EXTERNAL>
        if (a->adapt) { if (ext) AMCMC_LC(arwmh_block_kernel, true, true); else AMCMC_LC(arwmh_block_kernel, true, false); }
        else          { if (ext) AMCMC_LC(arwmh_block_kernel, false, true); else AMCMC_LC(arwmh_block_kernel, false, false); }
      }
#undef AMCMC_LC
#undef AMCMC_LC1
      return rc;
    }
  }
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
