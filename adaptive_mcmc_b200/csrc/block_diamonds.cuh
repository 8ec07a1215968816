// block_diamonds.cuh -- shared by block_diamonds.cu and its two run translation units (the fused-run kernels of the
// float32 and the float64 state are compiled separately: each is ~40 kernel instantiations, the long pole of the build).
#pragma once
#include "launch_block.cuh"

namespace amcmc {

template <typename R> inline DiamondsBlockModel<R> make_dm(const amcmc_model* m) {
  DiamondsBlockModel<R> b;
  b.d = m->dim;
  b.kc = m->dim - 2;
  b.n = (int)m->n_rows;
  b.n_stride = (int)m->arr_len[0];
  b.XcT = (const R*)m->d_arr[0];
  b.Y = (const R*)m->d_arr[1];
  b.cst = m->cst;
  return b;
}

int run_diamonds_block_f32(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s);
int run_diamonds_block_f64(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s);

}  // namespace amcmc
