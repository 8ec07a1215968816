// small_std_normal.cu -- N(0, I_d) targets (d in {1, 2, 4, 8}) on the thread-per-chain register kernel.
// Target of the reference's many-chain experiments: python/jupyter/asumptions_check.ipynb cells 17-28.
#include "launch_small.cuh"

namespace amcmc {

#define AMCMC_SN_DISPATCH(CALL)                                               \
  switch (m->dim) {                                                           \
    case 1: { constexpr int DD = 1; CALL; }                                   \
    case 2: { constexpr int DD = 2; CALL; }                                   \
    case 4: { constexpr int DD = 4; CALL; }                                   \
    case 8: { constexpr int DD = 8; CALL; }                                   \
    default:                                                                  \
      set_error("std_normal: thread-per-chain kernel compiled for d in {1,2,4,8}, got %d", m->dim); \
      return AMCMC_ERR_UNSUPPORTED;                                           \
  }

int run_std_normal(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  if (m->dtype == AMCMC_F32) {
    AMCMC_SN_DISPATCH(return (launch_small_run<StdNormalModel<float, DD>, float>(StdNormalModel<float, DD>{}, st, a, s)))
  }
  AMCMC_SN_DISPATCH(return (launch_small_run<StdNormalModel<double, DD>, double>(StdNormalModel<double, DD>{}, st, a, s)))
}

int init_std_normal(const amcmc_model* m, const amcmc_state* st, uint64_t seed, int64_t chain_offset, double radius,
                    int use_given_z, cudaStream_t s) {
  if (m->dtype == AMCMC_F32) {
    AMCMC_SN_DISPATCH(return (launch_small_init<StdNormalModel<float, DD>, float>(StdNormalModel<float, DD>{}, st, seed, chain_offset, radius, use_given_z, s)))
  }
  AMCMC_SN_DISPATCH(return (launch_small_init<StdNormalModel<double, DD>, double>(StdNormalModel<double, DD>{}, st, seed, chain_offset, radius, use_given_z, s)))
}

int potential_std_normal(const amcmc_model* m, int64_t n, const void* q, void* out, cudaStream_t s) {
  if (m->dtype == AMCMC_F32) {
    AMCMC_SN_DISPATCH(return (launch_small_potential<StdNormalModel<float, DD>, float>(StdNormalModel<float, DD>{}, n, q, out, s)))
  }
  AMCMC_SN_DISPATCH(return (launch_small_potential<StdNormalModel<double, DD>, double>(StdNormalModel<double, DD>{}, n, q, out, s)))
}

}  // namespace amcmc
