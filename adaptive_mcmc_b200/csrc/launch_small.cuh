// launch_small.cuh -- host-side launch helpers for the thread-per-chain kernels.
#pragma once
#include "arwmh_small.cuh"
#include "asss_small.cuh"
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
