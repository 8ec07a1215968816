// pooled.cu -- pooled (cross-chain) adaptation: sufficient statistics, Robbins-Monro update with an
// on-device Cholesky, and the frozen-kernel run entry that routes diamonds to the tensor cores.
// Not in the reference; spec in include/amcmc.h and DESIGN.md (BASELINE.json configs[3]).
#include "internal.h"
#include "common.cuh"

namespace amcmc {

constexpr int kStatsChains = 128;  // chains per CTA in the statistics kernel

// out[0] += chains; out[1+k] += sum delta_k; out[1+d+e] += sum delta_i delta_j; out[1+d+np] += sum macc
template <typename R>
__global__ void __launch_bounds__(256) pooled_stats_kernel(int64_t C, int d, const R* __restrict__ z,
                                                           const R* __restrict__ loc, const R* __restrict__ macc,
                                                           double* __restrict__ out) {
  extern __shared__ float sdelta[];  // [d][kStatsChains + 1] as float/double? keep as double for fp64 states
  double* dl = reinterpret_cast<double*>(sdelta);
  const int64_t c0 = (int64_t)blockIdx.x * kStatsChains;
  const int nc = (int)min((int64_t)kStatsChains, C - c0);
  const int ld = kStatsChains + 1;
  for (int idx = threadIdx.x; idx < d * kStatsChains; idx += blockDim.x) {
    const int k = idx / kStatsChains, cc = idx % kStatsChains;
    dl[k * ld + cc] = (cc < nc) ? (double)z[(int64_t)k * C + c0 + cc] - (double)loc[k] : 0.0;
  }
  __syncthreads();
  const int np = d * (d + 1) / 2;
  for (int e = threadIdx.x; e < np + d + 2; e += blockDim.x) {
    double s = 0;
    if (e < d) {
      for (int cc = 0; cc < nc; ++cc) s += dl[e * ld + cc];
      atomicAdd(&out[1 + e], s);
    } else if (e < d + np) {
      const int t = e - d;
      int i = (int)((sqrtf(1.0f + 8.0f * (float)t) - 1.0f) * 0.5f);
      while (i * (i + 1) / 2 > t) --i;
      while ((i + 1) * (i + 2) / 2 <= t) ++i;
      const int j = t - i * (i + 1) / 2;
      for (int cc = 0; cc < nc; ++cc) s += dl[i * ld + cc] * dl[j * ld + cc];
      atomicAdd(&out[1 + d + t], s);
    } else if (e == d + np) {
      for (int cc = 0; cc < nc; ++cc) s += (double)macc[c0 + cc];
      atomicAdd(&out[1 + d + np], s);
    } else {
      atomicAdd(&out[0], (double)nc);
    }
  }
}

// one CTA: Robbins-Monro + Cholesky in float64
template <typename R>
__global__ void __launch_bounds__(256) pooled_update_kernel(int d, const double* __restrict__ st, R* __restrict__ loc,
                                                            R* __restrict__ scale, R* __restrict__ lam,
                                                            double* __restrict__ cov, double gamma, double target) {
  extern __shared__ double sm[];
  double* A = sm;            // [d*d] working copy (lower)
  __shared__ int ok;
  const int np = d * (d + 1) / 2;
  const double n = st[0];
  const double inv = n > 0 ? 1.0 / n : 0.0;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    const int a = i >= j ? i : j, b = i >= j ? j : i;
    const double S = st[1 + d + a * (a + 1) / 2 + b] * inv;
    const double cnew = (1.0 - gamma) * cov[e] + gamma * S;
    cov[e] = cnew;
    A[e] = cnew;
  }
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += blockDim.x) loc[k] = (R)((double)loc[k] + gamma * st[1 + k] * inv);
  if (threadIdx.x == 0) lam[0] = (R)((double)lam[0] + gamma * (st[1 + d + np] * inv - target));
  // right-looking Cholesky, column by column
  for (int j = 0; j < d; ++j) {
    __syncthreads();
    const double piv = A[j * d + j];
    if (!(piv > 0.0) || !(piv < 1e300)) {
      if (threadIdx.x == 0) ok = 0;
      break;
    }
    const double r = sqrt(piv);
    __syncthreads();
    for (int i = j + threadIdx.x; i < d; i += blockDim.x) A[i * d + j] = (i == j) ? r : A[i * d + j] / r;
    __syncthreads();
    for (int e = threadIdx.x; e < (d - j - 1) * (d - j - 1); e += blockDim.x) {
      const int i = j + 1 + e / (d - j - 1), k = j + 1 + e % (d - j - 1);
      if (k <= i) A[i * d + k] -= A[i * d + j] * A[k * d + j];
    }
  }
  __syncthreads();
  if (ok)
    for (int e = threadIdx.x; e < np; e += blockDim.x) {
      int i = (int)((sqrtf(1.0f + 8.0f * (float)e) - 1.0f) * 0.5f);
      while (i * (i + 1) / 2 > e) --i;
      while ((i + 1) * (i + 2) / 2 <= e) ++i;
      const int j = e - i * (i + 1) / 2;
      scale[e] = (R)A[i * d + j];
    }
}

// broadcast the shared adaptation state into the per-chain arrays (generic CUDA-core path)
template <typename R>
__global__ void pooled_broadcast_kernel(int64_t C, int d, const R* __restrict__ loc, const R* __restrict__ scale,
                                        const R* __restrict__ lam, R* __restrict__ cl, R* __restrict__ cs,
                                        R* __restrict__ clam) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int np = d * (d + 1) / 2;
  for (int k = 0; k < d; ++k) cl[(int64_t)k * C + c] = loc[k];
  for (int e = 0; e < np; ++e) cs[(int64_t)e * C + c] = scale[e];
  clam[c] = lam[0];
}

// Every frozen kernel (thread-per-chain, CTA-per-chain, tensor-core) reports mean_accept_prob as the mean acceptance
// probability over exactly this call; amcmc_pooled_stats averages it over the chains for the Robbins-Monro step.

bool diamonds_tc_available(const amcmc_model* m);
int run_diamonds_tc(const amcmc_model* m, const amcmc_state* st, const void* loc, const void* scale_packed,
                    const void* log_step, const amcmc_run_args* a, cudaStream_t s);

}  // namespace amcmc

using namespace amcmc;

extern "C" {

int amcmc_pooled_run(const amcmc_model* m, amcmc_state* st, const amcmc_pooled* pool, const amcmc_run_args* args,
                     void* stream) {
  if (!m || !st || !pool || !args) { set_error("amcmc_pooled_run: NULL argument"); return AMCMC_ERR_ARG; }
  if (pool->dim != m->dim || pool->dtype != m->dtype || st->dim != m->dim || st->dtype != m->dtype) {
    set_error("amcmc_pooled_run: dim/dtype mismatch between model, state and pool");
    return AMCMC_ERR_ARG;
  }
  if (!pool->loc || !pool->scale || !pool->log_step_size) { set_error("amcmc_pooled_run: NULL pool array"); return AMCMC_ERR_ARG; }
  if (args->n_steps < 0 || args->thinning < 1 || args->collect_start < 0) {
    set_error("amcmc_pooled_run: need n_steps >= 0, thinning >= 1, collect_start >= 0");
    return AMCMC_ERR_ARG;
  }
  if (args->n_steps == 0) return AMCMC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  const bool want_tc = (args->impl == 0 || args->impl == 3) && diamonds_tc_available(m);
  if (args->impl == 3 && !want_tc) { set_error("amcmc_pooled_run: tensor-core path unavailable for this model/dtype"); return AMCMC_ERR_UNSUPPORTED; }
  if (want_tc) {
    if (args->rng_mode == AMCMC_RNG_EXTERNAL && (!args->normals || !args->uniforms)) {
      set_error("amcmc_pooled_run: rng_mode EXTERNAL needs normals and uniforms");
      return AMCMC_ERR_ARG;
    }
    rc = run_diamonds_tc(m, st, pool->loc, pool->scale, pool->log_step_size, args, s);
    if (rc == AMCMC_OK) st->i += args->n_steps;
    return rc;
  }
  // generic: broadcast, then the frozen CUDA-core kernels
  const unsigned grid = (unsigned)((st->n_chains + 127) / 128);
  if (m->dtype == AMCMC_F32)
    pooled_broadcast_kernel<float><<<grid, 128, 0, s>>>(st->n_chains, m->dim, (const float*)pool->loc, (const float*)pool->scale,
                                                        (const float*)pool->log_step_size, (float*)st->loc, (float*)st->scale,
                                                        (float*)st->log_step_size);
  else
    pooled_broadcast_kernel<double><<<grid, 128, 0, s>>>(st->n_chains, m->dim, (const double*)pool->loc, (const double*)pool->scale,
                                                         (const double*)pool->log_step_size, (double*)st->loc, (double*)st->scale,
                                                         (double*)st->log_step_size);
  if ((rc = check_cuda(cudaGetLastError(), "pooled_broadcast_kernel launch"))) return rc;
  amcmc_run_args a = *args;
  a.adapt = 0;
  a.kernel_kind = AMCMC_KERNEL_ARWMH;
  return amcmc_arwmh_run(m, st, &a, stream);
}

int amcmc_pooled_stats(const amcmc_state* st, const amcmc_pooled* pool, double* out, void* stream) {
  if (!st || !pool || !out) { set_error("amcmc_pooled_stats: NULL argument"); return AMCMC_ERR_ARG; }
  const int d = st->dim;
  if (d > 64) { set_error("amcmc_pooled_stats: d <= 64 supported (got %d)", d); return AMCMC_ERR_UNSUPPORTED; }
  cudaStream_t s = (cudaStream_t)stream;
  const int n = 2 + d + d * (d + 1) / 2;
  int rc = check_cuda(cudaMemsetAsync(out, 0, sizeof(double) * n, s), "cudaMemsetAsync");
  if (rc) return rc;
  const unsigned grid = (unsigned)((st->n_chains + kStatsChains - 1) / kStatsChains);
  const size_t smem = sizeof(double) * d * (kStatsChains + 1);
  if (st->dtype == AMCMC_F32) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(pooled_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pooled_stats_kernel<float><<<grid, 256, smem, s>>>(st->n_chains, d, (const float*)st->z, (const float*)pool->loc,
                                                       (const float*)st->mean_accept_prob, out);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(pooled_stats_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pooled_stats_kernel<double><<<grid, 256, smem, s>>>(st->n_chains, d, (const double*)st->z, (const double*)pool->loc,
                                                        (const double*)st->mean_accept_prob, out);
  }
  return check_cuda(cudaGetLastError(), "pooled_stats_kernel launch");
}

int amcmc_pooled_update(amcmc_pooled* pool, const double* stats, double lr_decay, double target, void* stream) {
  if (!pool || !stats || !pool->cov) { set_error("amcmc_pooled_update: NULL argument"); return AMCMC_ERR_ARG; }
  const int d = pool->dim;
  if (d > 64) { set_error("amcmc_pooled_update: d <= 64 supported (got %d)", d); return AMCMC_ERR_UNSUPPORTED; }
  cudaStream_t s = (cudaStream_t)stream;
  const double n = (double)(pool->window + 1);
  const double gamma = 1.0 / pow(n, lr_decay);
  const size_t smem = sizeof(double) * d * d;
  if (pool->dtype == AMCMC_F32)
    pooled_update_kernel<float><<<1, 256, smem, s>>>(d, stats, (float*)pool->loc, (float*)pool->scale,
                                                     (float*)pool->log_step_size, pool->cov, gamma, target);
  else
    pooled_update_kernel<double><<<1, 256, smem, s>>>(d, stats, (double*)pool->loc, (double*)pool->scale,
                                                      (double*)pool->log_step_size, pool->cov, gamma, target);
  int rc = check_cuda(cudaGetLastError(), "pooled_update_kernel launch");
  if (rc == AMCMC_OK) pool->window += 1;
  return rc;
}

}  // extern "C"
