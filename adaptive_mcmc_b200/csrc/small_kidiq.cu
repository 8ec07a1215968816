// small_kidiq.cu -- kidiq (d = 4, N = 434 rows) on the thread-per-chain register kernel.
// Model: python/scripts/run_kidiq_kidscore_lr_decay.py:29-41.
#include "launch_small.cuh"

namespace amcmc {

template <typename R> static KidiqModel<R> make_kidiq(const amcmc_model* m) {
  KidiqModel<R> k;
  k.kid = (const R*)m->d_arr[0];
  k.hs = (const R*)m->d_arr[1];
  k.iq = (const R*)m->d_arr[2];
  k.n = (int)m->n_rows;
  k.cst = (R)m->cst;
  return k;
}

int run_kidiq(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  if (m->dtype == AMCMC_F32) return launch_small_run<KidiqModel<float>, float>(make_kidiq<float>(m), st, a, s);
  return launch_small_run<KidiqModel<double>, double>(make_kidiq<double>(m), st, a, s);
}

int init_kidiq(const amcmc_model* m, const amcmc_state* st, uint64_t seed, int64_t chain_offset, double radius,
               int use_given_z, cudaStream_t s) {
  if (m->dtype == AMCMC_F32)
    return launch_small_init<KidiqModel<float>, float>(make_kidiq<float>(m), st, seed, chain_offset, radius, use_given_z, s);
  return launch_small_init<KidiqModel<double>, double>(make_kidiq<double>(m), st, seed, chain_offset, radius, use_given_z, s);
}

int potential_kidiq(const amcmc_model* m, int64_t n, const void* q, void* out, cudaStream_t s) {
  if (m->dtype == AMCMC_F32) return launch_small_potential<KidiqModel<float>, float>(make_kidiq<float>(m), n, q, out, s);
  return launch_small_potential<KidiqModel<double>, double>(make_kidiq<double>(m), n, q, out, s);
}

}  // namespace amcmc
