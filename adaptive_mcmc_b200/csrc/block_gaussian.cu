// block_gaussian.cu -- correlated Gaussian target N(0, Sigma), Sigma^-1 = P P^T (BASELINE.json configs[4]):
// ARWMH on the block kernel for d <= 32 and the RAM variant (ram_block.cuh) for d <= 256.
#include <cmath>
#include <vector>
#include "launch_block.cuh"
#include "ram_block.cuh"

namespace amcmc {

// P: host row-major dense lower Cholesky factor of the precision.  Uploads the dense factor and, when
// the factor is banded (e.g. the AR(1) target: bidiagonal), the band in [bw+1][d] layout.
int create_gaussian(amcmc_model* m, const double* P, int d) {
  // bandwidth up to round-off: entries below 1e-13 of the largest are treated as structural zeros
  double pmax = 0;
  for (int i = 0; i < d; ++i)
    for (int j = 0; j <= i; ++j) pmax = std::fmax(pmax, std::fabs(P[(size_t)i * d + j]));
  int bw = 0;
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < i; ++j)
      if (std::fabs(P[(size_t)i * d + j]) > 1e-13 * pmax && i - j > bw) bw = i - j;
  std::vector<double> dense((size_t)d * d, 0.0), band((size_t)(bw + 1) * d, 0.0);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j <= i; ++j) {
      dense[(size_t)i * d + j] = P[(size_t)i * d + j];
      if (i - j <= bw) band[(size_t)(i - j) * d + j] = P[(size_t)i * d + j];
    }
  m->n_rows = bw;  // bandwidth
  const size_t w = m->dtype == AMCMC_F64 ? 8 : 4;
  int rc;
  auto up = [&](const std::vector<double>& src, void** dst) -> int {
    if ((rc = check_cuda(cudaMalloc(dst, src.size() * w), "cudaMalloc(P)"))) return rc;
    if (m->dtype == AMCMC_F64) return check_cuda(cudaMemcpy(*dst, src.data(), src.size() * 8, cudaMemcpyHostToDevice), "cudaMemcpy");
    std::vector<float> f(src.begin(), src.end());
    return check_cuda(cudaMemcpy(*dst, f.data(), f.size() * 4, cudaMemcpyHostToDevice), "cudaMemcpy");
  };
  if ((rc = up(dense, &m->d_arr[0]))) return rc;
  if ((rc = up(band, &m->d_arr[1]))) return rc;
  m->cst = 0;
  return AMCMC_OK;
}

static bool use_band(const amcmc_model* m) { return m->n_rows <= 8; }

template <typename R> static GaussianBlockModel<R> make_dense(const amcmc_model* m) {
  GaussianBlockModel<R> g;
  g.d = m->dim;
  g.P = (const R*)m->d_arr[0];
  return g;
}
template <typename R> static GaussianBandModel<R> make_band(const amcmc_model* m) {
  GaussianBandModel<R> g;
  g.d = m->dim;
  g.bw = (int)m->n_rows;
  g.Pband = (const R*)m->d_arr[1];
  return g;
}

template <class BM, typename R>
static int launch_ram(const BM& bm, int d, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  const StateView<R> sv = make_state_view<R>(st);
  const RunView<R> rv = make_run_view<R>(st, a);
  const size_t smem = RamSmem<R>::bytes(d);
  if (smem > 227 * 1024) {
    set_error("RAM kernel: d = %d needs %zu bytes of shared memory (> 227 KB)", d, smem);
    return AMCMC_ERR_UNSUPPORTED;
  }
  int rc;
  if (a->rng_mode == AMCMC_RNG_EXTERNAL) {
    auto k = ram_block_kernel<BM, R, true>;
    if ((rc = ensure_smem(k, smem))) return rc;
    k<<<(unsigned)st->n_chains, kRamThreads, smem, s>>>(bm, sv, rv, d);
  } else {
    auto k = ram_block_kernel<BM, R, false>;
    if ((rc = ensure_smem(k, smem))) return rc;
    k<<<(unsigned)st->n_chains, kRamThreads, smem, s>>>(bm, sv, rv, d);
  }
  return check_cuda(cudaGetLastError(), "ram_block_kernel launch");
}

int run_gaussian(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s) {
  const int d = m->dim;
  if (a->kernel_kind == AMCMC_KERNEL_RAM) {
    if (d > 256) { set_error("RAM kernel: d <= 256 (got %d)", d); return AMCMC_ERR_UNSUPPORTED; }
    if (!a->adapt) { set_error("RAM kernel: frozen mode not available"); return AMCMC_ERR_UNSUPPORTED; }
    if (m->dtype == AMCMC_F32)
      return use_band(m) ? launch_ram<GaussianBandModel<float>, float>(make_band<float>(m), d, st, a, s)
                         : launch_ram<GaussianBlockModel<float>, float>(make_dense<float>(m), d, st, a, s);
    return use_band(m) ? launch_ram<GaussianBandModel<double>, double>(make_band<double>(m), d, st, a, s)
                       : launch_ram<GaussianBlockModel<double>, double>(make_dense<double>(m), d, st, a, s);
  }
  if (m->dtype == AMCMC_F32) return launch_block_run<GaussianBlockModel<float>, float>(make_dense<float>(m), d, st, a, s);
  return launch_block_run<GaussianBlockModel<double>, double>(make_dense<double>(m), d, st, a, s);
}

int init_gaussian(const amcmc_model* m, const amcmc_state* st, uint64_t seed, int64_t chain_offset, double radius,
                  int use_given_z, cudaStream_t s) {
  if (m->dtype == AMCMC_F32)
    return launch_block_init<GaussianBlockModel<float>, float>(make_dense<float>(m), m->dim, st, seed, chain_offset, radius, use_given_z, s);
  return launch_block_init<GaussianBlockModel<double>, double>(make_dense<double>(m), m->dim, st, seed, chain_offset, radius, use_given_z, s);
}

int potential_gaussian(const amcmc_model* m, int64_t n, const void* q, void* out, cudaStream_t s) {
  if (m->dtype == AMCMC_F32)
    return launch_block_potential<GaussianBlockModel<float>, float>(make_dense<float>(m), m->dim, n, q, out, s);
  return launch_block_potential<GaussianBlockModel<double>, double>(make_dense<double>(m), m->dim, n, q, out, s);
}

}  // namespace amcmc
