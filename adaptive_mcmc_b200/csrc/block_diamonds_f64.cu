// block_diamonds_f64.cu -- the fused-run kernels of the diamonds block path for double state (see block_diamonds.cuh)
#include "block_diamonds.cuh"

namespace amcmc {

int run_diamonds_block_f64(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  return launch_block_run<DiamondsBlockModel<double>, double>(make_dm<double>(m), m->dim, st, a, s);
}

}  // namespace amcmc
