// small_eight_schools.cu -- eight_schools (d = 10) on the thread-per-chain register kernel.
// Model: python/scripts/run_eight_schools_lr_decay.py:26-35.
#include "launch_small.cuh"

namespace amcmc {

template <typename R> static EightSchoolsModel<R> make_es(const amcmc_model* m) {
  EightSchoolsModel<R> e;
  for (int j = 0; j < 8; ++j) {
    e.y[j] = (R)m->h_small[j];
    e.inv_sigma[j] = (R)(1.0 / m->h_small[8 + j]);
  }
  e.cst = (R)m->cst;
  return e;
}

int run_eight_schools(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  if (m->dtype == AMCMC_F32) return launch_small_run<EightSchoolsModel<float>, float>(make_es<float>(m), st, a, s);
  return launch_small_run<EightSchoolsModel<double>, double>(make_es<double>(m), st, a, s);
}

int init_eight_schools(const amcmc_model* m, const amcmc_state* st, uint64_t seed, int64_t chain_offset,
                       double radius, int use_given_z, cudaStream_t s) {
  if (m->dtype == AMCMC_F32)
    return launch_small_init<EightSchoolsModel<float>, float>(make_es<float>(m), st, seed, chain_offset, radius, use_given_z, s);
  return launch_small_init<EightSchoolsModel<double>, double>(make_es<double>(m), st, seed, chain_offset, radius, use_given_z, s);
}

int potential_eight_schools(const amcmc_model* m, int64_t n, const void* q, void* out, cudaStream_t s) {
  if (m->dtype == AMCMC_F32) return launch_small_potential<EightSchoolsModel<float>, float>(make_es<float>(m), n, q, out, s);
  return launch_small_potential<EightSchoolsModel<double>, double>(make_es<double>(m), n, q, out, s);
}

}  // namespace amcmc
