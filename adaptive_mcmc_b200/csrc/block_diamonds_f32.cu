// block_diamonds_f32.cu -- the fused-run kernels of the diamonds block path for float state (see block_diamonds.cuh)
#include "block_diamonds.cuh"

namespace amcmc {

int run_diamonds_block_f32(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  return launch_block_run<DiamondsBlockModel<float>, float>(make_dm<float>(m), m->dim, st, a, s);
}

}  // namespace amcmc
