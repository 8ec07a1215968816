// launch_small.cuh -- host-side launch helpers for the thread-per-chain kernels.
#pragma once
#include "arwmh_small.cuh"
#include "asss_small.cuh"
#include <cstdlib>
#include "internal.h"

namespace amcmc {

template <typename R> inline StateView<R> make_state_view(const amcmc_state* st) {
  StateView<R> v;
  v.C = st->n_chains;
  v.z = (R*)st->z;
  v.pe = (R*)st->potential_energy;
  v.macc = (R*)st->mean_accept_prob;
  v.loc = (R*)st->loc;
  v.scale = (R*)st->scale;
  v.lam = (R*)st->log_step_size;
  v.asc = (R*)st->as_change;
  return v;
}

template <typename R> inline RunView<R> make_run_view(const amcmc_state* st, const amcmc_run_args* a) {
  RunView<R> r;
  r.i0 = st->i;
  r.n_steps = a->n_steps;
  r.thinning = a->thinning;
  r.collect_start = a->collect_start;
  r.num_warmup = a->num_warmup;
  r.lr_decay = (R)a->lr_decay;
  r.target = (R)a->target_accept_prob;
  r.eps = (R)a->eps;
  r.seed = a->seed;
  r.chain_offset = a->chain_offset;
  r.normals = (const R*)a->normals;
  r.uniforms = (const R*)a->uniforms;
  r.out_z = (R*)a->out_z;
  r.out_pe = (R*)a->out_potential_energy;
  r.out_acc = a->out_accept;
  return r;
}

// Thread-per-chain launch.  65,536 chains / 64 threads = 1024 CTAs = 6.9 CTAs per SM on 148 SMs:
// a single wave at 7 resident CTAs (launch bounds cap the kernel at 144 registers for that).
template <class Model, typename R>
int launch_small_run(const Model& m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  const StateView<R> sv = make_state_view<R>(st);
  const RunView<R> rv = make_run_view<R>(st, a);
  const int block = 64;
  const unsigned grid = (unsigned)((st->n_chains + block - 1) / block);
  const bool ext = a->rng_mode == AMCMC_RNG_EXTERNAL;
  // Few chains (up to 6 groups of 32 per SM): a producer warp per 32 chains generates the draws one step ahead
  // (arwmh_small_duo_kernel).  AMCMC_SMALL_DUO=0/1 overrides (read per launch: tests switch it).
  if (!ext && a->impl != 4 && a->n_steps > 0) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const char* e = getenv("AMCMC_SMALL_DUO");
    const int64_t n_groups = (st->n_chains + 31) / 32;
    // measured at 10,000 fused steps, chains -> ms (one warp / two warps per group): 4,736 -> 7.19 / 4.74, 9,472 -> 7.19 / 4.79,
    // 18,944 -> 7.26 / 7.01, 28,416 -> 10.45 / 9.99, 37,888 -> 10.50 / 13.99: up to 6 groups per SM
    const bool duo = e ? atoi(e) != 0 : (n_groups <= (int64_t)6 * sms && a->n_steps >= 32);
    if (duo) {
      const unsigned dgrid = (unsigned)n_groups;
      if (a->kernel_kind == AMCMC_KERNEL_ASSS) {
        if (a->adapt) asss_small_duo_kernel<Model, R, true><<<dgrid, 64, 0, s>>>(m, sv, rv);
        else asss_small_duo_kernel<Model, R, false><<<dgrid, 64, 0, s>>>(m, sv, rv);
      } else {
        if (a->adapt) arwmh_small_duo_kernel<Model, R, true><<<dgrid, 64, 0, s>>>(m, sv, rv);
        else arwmh_small_duo_kernel<Model, R, false><<<dgrid, 64, 0, s>>>(m, sv, rv);
      }
      return check_cuda(cudaGetLastError(), "small duo kernel launch");
    }
  }
  // Balanced variant (arwmh_small.cuh): when the chains give every scheduler more than ~2 warps but not a whole number of
  // them, one CTA of 12 worker warps per SM with a work queue keeps all schedulers saturated.  AMCMC_SMALL_BALANCED=0 forces the plain
  // kernel, =1 the balanced one wherever it fits (tests run both).
  if constexpr (sizeof(R) == 4) {
    static const int forced = [] { const char* e = getenv("AMCMC_SMALL_BALANCED"); return e ? atoi(e) : -1; }();
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_groups = (st->n_chains + 31) / 32;
    const int64_t per_cta = (n_groups + sms - 1) / sms;
    const size_t slot_bytes = (size_t)ChainSlot<R, Model::D>::NREG * 32 * sizeof(R);
    const size_t smem = (size_t)per_cta * slot_bytes;
    const bool fits = sizeof(R) == 4 && per_cta <= kBalQueue && smem <= 200 * 1024 && a->n_steps > 0;
    // auto: launches long enough to amortise the hand-offs, and a group count per SM that the plain kernel cannot spread evenly
    const bool want = a->impl == 4 ? true : a->impl == 1 ? false : forced >= 0 ? forced != 0
                      : (per_cta >= 8 && per_cta % 4 != 0 && a->n_steps >= 128);
    if (a->impl == 4 && !fits) {
      set_error("amcmc_arwmh_run: the balanced thread-per-chain kernel needs fp32 and at most %d groups of 32 chains per SM within 200 KB of shared memory", kBalQueue);
      return AMCMC_ERR_UNSUPPORTED;
    }
    if (fits && want) {
      const unsigned bgrid = (unsigned)(n_groups < sms ? n_groups : sms);
      int64_t seg64 = a->n_steps / 64;  // ~64 hand-offs per group and launch: the tail imbalance is ~1 segment
      if (seg64 < 16) seg64 = 16;
      if (seg64 > 512) seg64 = 512;
      static const int seg_env = [] { const char* e = getenv("AMCMC_SMALL_BALANCED_SEG"); return e ? atoi(e) : 0; }();  // tuning knob
      const int seg = seg_env > 0 ? seg_env : (int)seg64;
      int rc = 0;
      int workers = kBalWarps;
      auto go = [&](auto kern) {
        if (smem > 48 * 1024) rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(balanced)");
        if (!rc) kern<<<bgrid, 32 * workers, smem, s>>>(m, sv, rv, n_groups, seg);
      };
      if (a->kernel_kind == AMCMC_KERNEL_ASSS) {
        workers = kBalWarpsAsss;  // 8 workers = 2 per scheduler: 255 registers per thread, the slice-sampling step does not spill
        if (a->adapt) { if (ext) go(arwmh_small_balanced_kernel<Model, R, true, AsssRange<Model, R, true, true>, kBalWarpsAsss>); else go(arwmh_small_balanced_kernel<Model, R, true, AsssRange<Model, R, true, false>, kBalWarpsAsss>); }
        else { if (ext) go(arwmh_small_balanced_kernel<Model, R, false, AsssRange<Model, R, false, true>, kBalWarpsAsss>); else go(arwmh_small_balanced_kernel<Model, R, false, AsssRange<Model, R, false, false>, kBalWarpsAsss>); }
      } else {
        if (a->adapt) { if (ext) go(arwmh_small_balanced_kernel<Model, R, true, ArwmhRange<Model, R, true, true>>); else go(arwmh_small_balanced_kernel<Model, R, true, ArwmhRange<Model, R, true, false>>); }
        else { if (ext) go(arwmh_small_balanced_kernel<Model, R, false, ArwmhRange<Model, R, false, true>>); else go(arwmh_small_balanced_kernel<Model, R, false, ArwmhRange<Model, R, false, false>>); }
      }
      if (rc) return rc;
      return check_cuda(cudaGetLastError(), "arwmh_small_balanced_kernel launch");
    }
  }
  if (a->impl == 4) {
    set_error("amcmc_arwmh_run: the balanced thread-per-chain kernel is built for fp32");
    return AMCMC_ERR_UNSUPPORTED;
  }
  if (a->kernel_kind == AMCMC_KERNEL_ASSS) {
    if (a->adapt) {
      if (ext) asss_small_kernel<Model, R, true, true><<<grid, block, 0, s>>>(m, sv, rv);
      else     asss_small_kernel<Model, R, false, true><<<grid, block, 0, s>>>(m, sv, rv);
    } else {  // frozen adapt_state: ASSS.sample_Pnx (asss.py:279-315)
      if (ext) asss_small_kernel<Model, R, true, false><<<grid, block, 0, s>>>(m, sv, rv);
      else     asss_small_kernel<Model, R, false, false><<<grid, block, 0, s>>>(m, sv, rv);
    }
    return check_cuda(cudaGetLastError(), "asss_small_kernel launch");
  }
  if (a->adapt) {
    if (ext) arwmh_small_kernel<Model, R, true, true><<<grid, block, 0, s>>>(m, sv, rv);
    else     arwmh_small_kernel<Model, R, true, false><<<grid, block, 0, s>>>(m, sv, rv);
  } else {
    if (ext) arwmh_small_kernel<Model, R, false, true><<<grid, block, 0, s>>>(m, sv, rv);
    else     arwmh_small_kernel<Model, R, false, false><<<grid, block, 0, s>>>(m, sv, rv);
  }
  return check_cuda(cudaGetLastError(), "arwmh_small_kernel launch");
}

template <class Model, typename R>
int launch_small_init(const Model& m, const amcmc_state* st, uint64_t seed, int64_t chain_offset, double radius,
                      int use_given_z, cudaStream_t s) {
  const StateView<R> sv = make_state_view<R>(st);
  const int block = 128;
  const unsigned grid = (unsigned)((st->n_chains + block - 1) / block);
  arwmh_small_init_kernel<Model, R><<<grid, block, 0, s>>>(m, sv, seed, chain_offset, (R)radius, use_given_z);
  return check_cuda(cudaGetLastError(), "arwmh_small_init_kernel launch");
}

template <class Model, typename R>
int launch_small_potential(const Model& m, int64_t n, const void* q, void* out, cudaStream_t s) {
  const int block = 128;
  const unsigned grid = (unsigned)((n + block - 1) / block);
  potential_small_kernel<Model, R><<<grid, block, 0, s>>>(m, n, (const R*)q, (R*)out);
  return check_cuda(cudaGetLastError(), "potential_small_kernel launch");
}

}  // namespace amcmc
