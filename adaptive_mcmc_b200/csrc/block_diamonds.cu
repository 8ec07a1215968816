// block_diamonds.cu -- diamonds (d = K + 1 = 26, N = 5000) on the block-per-chain CUDA-core kernel:
// the exact fp64/fp32 parity path and the few-chain path (BASELINE.json configs[1]).
// Model: python/scripts/run_diamonds_lr_decay.py:24-40.
#include <cmath>
#include <vector>
#include "block_diamonds.cuh"

namespace amcmc {

// Host-side model preparation: centre the predictors in float64 (run_diamonds_lr_decay.py:28-29),
// transpose to [Kc][n_stride] so that consecutive threads read consecutive rows, upload.
int create_diamonds(amcmc_model* m, const double* X, int64_t n, int K, const double* Y) {
  const int kc = K - 1;
  const int64_t ns = (n + 31) & ~(int64_t)31;
  std::vector<double> xt((size_t)kc * ns, 0.0);
  for (int k = 0; k < kc; ++k) {
    double mean = 0;
    for (int64_t r = 0; r < n; ++r) mean += X[r * K + 1 + k];
    mean /= (double)n;
    for (int64_t r = 0; r < n; ++r) xt[(size_t)k * ns + r] = X[r * K + 1 + k] - mean;
  }
  m->n_rows = n;
  m->arr_len[0] = ns;  // leading dimension
  const size_t w = m->dtype == AMCMC_F64 ? 8 : 4;
  int rc;
  if ((rc = check_cuda(cudaMalloc(&m->d_arr[0], xt.size() * w), "cudaMalloc(XcT)"))) return rc;
  if ((rc = check_cuda(cudaMalloc(&m->d_arr[1], (size_t)n * w), "cudaMalloc(Y)"))) return rc;
  if (m->dtype == AMCMC_F64) {
    if ((rc = check_cuda(cudaMemcpy(m->d_arr[0], xt.data(), xt.size() * 8, cudaMemcpyHostToDevice), "cudaMemcpy"))) return rc;
    if ((rc = check_cuda(cudaMemcpy(m->d_arr[1], Y, (size_t)n * 8, cudaMemcpyHostToDevice), "cudaMemcpy"))) return rc;
  } else {
    std::vector<float> xf(xt.begin(), xt.end()), yf(Y, Y + n);
    if ((rc = check_cuda(cudaMemcpy(m->d_arr[0], xf.data(), xf.size() * 4, cudaMemcpyHostToDevice), "cudaMemcpy"))) return rc;
    if ((rc = check_cuda(cudaMemcpy(m->d_arr[1], yf.data(), (size_t)n * 4, cudaMemcpyHostToDevice), "cudaMemcpy"))) return rc;
  }
  // folded constants: Kc*1/2 log 2pi + 2*Z_t(3;10) - log 2 + N*1/2 log 2pi,
  // Z_t = log c + 1/2 log nu + 1/2 log pi + lgamma(nu/2) - lgamma((nu+1)/2)
  const double h = 0.91893853320467274178;
  const double Z = std::log(10.0) + 0.5 * std::log(3.0) + 0.5 * std::log(M_PI) + std::lgamma(1.5) - std::lgamma(2.0);
  m->cst = kc * h + 2.0 * Z - std::log(2.0) + (double)n * h;
  return AMCMC_OK;
}

int run_diamonds_block(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  return m->dtype == AMCMC_F32 ? run_diamonds_block_f32(m, st, a, s) : run_diamonds_block_f64(m, st, a, s);
}

int init_diamonds(const amcmc_model* m, const amcmc_state* st, uint64_t seed, int64_t chain_offset, double radius,
                  int use_given_z, cudaStream_t s) {
  if (m->dtype == AMCMC_F32)
    return launch_block_init<DiamondsBlockModel<float>, float>(make_dm<float>(m), m->dim, st, seed, chain_offset, radius, use_given_z, s);
  return launch_block_init<DiamondsBlockModel<double>, double>(make_dm<double>(m), m->dim, st, seed, chain_offset, radius, use_given_z, s);
}

int potential_diamonds_block(const amcmc_model* m, int64_t n, const void* q, void* out, cudaStream_t s) {
  if (m->dtype == AMCMC_F32)
    return launch_block_potential<DiamondsBlockModel<float>, float>(make_dm<float>(m), m->dim, n, q, out, s);
  return launch_block_potential<DiamondsBlockModel<double>, double>(make_dm<double>(m), m->dim, n, q, out, s);
}

}  // namespace amcmc
