// sinkhorn.cu -- entropy-regularised optimal transport between two uniform samples: the Sinkhorn iterations behind
// `wasserstein_sinkhorn` (reference: python/utils/evaluation.py:69-97, which delegates to OTT-JAX `linear.solve(geom)` on a
// PointCloud with the Euclidean cost and reads `ot.ent_reg_cost`).
//
// OTT-JAX is a third-party dependency that is not under /root/reference and not installable here; what is restated is its
// published algorithm with the defaults of `ott.solvers.linear.sinkhorn.Sinkhorn` as recalled (threshold 1e-3 on the L1 error
// of the row marginal, checked every 10 iterations, at most 2000 iterations, log-sum-exp mode, zero initial potentials, no
// momentum, one iteration = g-update then f-update; epsilon = 0.05 x mean cost when not given) and the definition of
// `ent_reg_cost` for a balanced problem: sum_i a_i f_i + sum_j b_j g_j + eps (1 - sum_ij P_ij) with
// P_ij = a_i b_j exp((f_i + g_j - C_ij) / eps), i.e. <P, C> + eps KL(P | a x b) at convergence.  For a converged run the value
// is the unique optimum of a strictly convex problem, so it does not depend on those details; parity: tests vs the float64
// NumPy restatement (oracle/evaluation_numpy.py) -- OTT itself is unpinned.
//
// Work per iteration: two log-sum-exp sweeps over the n x m float32 cost matrix (row-wise and column-wise), HBM-bound:
// 2 x 400 MB at 10^4 x 10^4.  The matrix is read with coalesced 128-byte warp accesses in both sweeps (the column sweep
// gives every warp 32 adjacent columns and strides over rows).
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "internal.h"

namespace amcmc {

struct Lse {  // running log-sum-exp: log(sum exp(v_k)) = m + log(s)
  float m, s;
};
__device__ __forceinline__ void lse_push(Lse& a, float v) {
  if (v > a.m) { a.s = a.s * __expf(a.m - v) + 1.f; a.m = v; }
  else a.s += __expf(v - a.m);
}
__device__ __forceinline__ void lse_merge(Lse& a, const Lse& b) {
  if (b.m > a.m) { a.s = a.s * __expf(a.m - b.m) + b.s; a.m = b.m; }
  else if (b.m > -FLT_MAX) a.s += b.s * __expf(b.m - a.m);
}

// f_i <- -eps LSE_j((g_j - C_ij) / eps + log b_j);  also accumulates the L1 error of the row marginal under (f_old, g):
// sum_j P_ij = a_i exp((f_old_i - f_new_i) / eps)
__global__ void __launch_bounds__(256) sinkhorn_rows_kernel(const float* __restrict__ C, int n, int m, const float* __restrict__ g,
                                                            float* __restrict__ f, float eps, float log_b, double* __restrict__ err,
                                                            double a_i) {
  const int i = blockIdx.x;
  const float inv = 1.f / eps;
  const float* row = C + (int64_t)i * m;
  Lse t{-FLT_MAX, 0.f};
  for (int j = threadIdx.x; j < m; j += 256) lse_push(t, (g[j] - row[j]) * inv);
  for (int o = 16; o; o >>= 1) {
    Lse u{__shfl_xor_sync(0xffffffffu, t.m, o), __shfl_xor_sync(0xffffffffu, t.s, o)};
    lse_merge(t, u);
  }
  __shared__ Lse sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) lse_merge(t, sh[w]);
    const float fn = -eps * (t.m + logf(t.s) + log_b);
    if (err) atomicAdd(err, a_i * fabs(exp(((double)f[i] - (double)fn) / (double)eps) - 1.0));
    f[i] = fn;
  }
}

// g_j <- -eps LSE_i((f_i - C_ij) / eps + log a_i): 32 adjacent columns per warp, 8 warps stride over the rows
__global__ void __launch_bounds__(256) sinkhorn_cols_kernel(const float* __restrict__ C, int n, int m, const float* __restrict__ f,
                                                            float* __restrict__ g, float eps, float log_a) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const float inv = 1.f / eps;
  Lse t{-FLT_MAX, 0.f};
  if (j < m)
    for (int i = w; i < n; i += 8) lse_push(t, (f[i] - C[(int64_t)i * m + j]) * inv);
  __shared__ Lse sh[8][33];
  sh[w][lane] = t;
  __syncthreads();
  if (w == 0 && j < m) {
    for (int k = 1; k < 8; ++k) lse_merge(t, sh[k][lane]);
    g[j] = -eps * (t.m + logf(t.s) + log_a);
  }
}

// sum_i a f_i, sum_j b g_j, sum_ij P_ij (for the non-converged correction)
__global__ void __launch_bounds__(256) sinkhorn_objective_kernel(const float* __restrict__ C, int n, int m, const float* __restrict__ f,
                                                                 const float* __restrict__ g, float eps, double a_i, double b_j,
                                                                 double* __restrict__ out) {
  const int i = blockIdx.x;
  const float* row = C + (int64_t)i * m;
  const float inv = 1.f / eps, fi = f[i];
  double s = 0.0;
  for (int j = threadIdx.x; j < m; j += 256) s += (double)__expf((fi + g[j] - row[j]) * inv);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) s += sh[w];
    atomicAdd(&out[0], a_i * (double)fi);
    atomicAdd(&out[2], a_i * b_j * s);
  }
  if (i == 0) {
    double sg = 0.0;
    for (int j = threadIdx.x; j < m; j += 256) sg += (double)g[j];
    for (int o = 16; o; o >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&out[1], b_j * sg);
  }
}

__global__ void sinkhorn_mean_kernel(const float* __restrict__ C, int64_t total, double* __restrict__ out) {
  double s = 0.0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) s += (double)C[k];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

}  // namespace amcmc

using namespace amcmc;

// cost: DEVICE float32 [n][m] (row-major).  epsilon <= 0: 0.05 x mean cost (OTT's default).  out_host[5]: ent_reg_cost,
// iterations run, last marginal error, converged (0/1), epsilon used.  f_out / g_out: optional DEVICE [n] / [m] potentials.
extern "C" int amcmc_eval_sinkhorn(const float* cost, int64_t n64, int64_t m64, double epsilon, double threshold, int max_iterations,
                                   int inner_iterations, float* f_out, float* g_out, double* out_host, void* stream) {
  if (!cost || !out_host || n64 < 1 || m64 < 1 || n64 > 200000 || m64 > 200000 || max_iterations < 1 || inner_iterations < 1) {
    set_error("amcmc_eval_sinkhorn: bad argument");
    return AMCMC_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)n64, m = (int)m64;
  int rc;
  char* buf = nullptr;
  const size_t bytes = ((size_t)(n + m) * 4 + 512 + 255) & ~(size_t)255;
  if ((rc = check_cuda(cudaMalloc(&buf, bytes), "cudaMalloc(sinkhorn)"))) return rc;
  double* scal = (double*)buf;  // [0] mean accumulator, [1] marginal error, [4..6] objective parts
  float* f = f_out ? f_out : (float*)(buf + 512);
  float* g = g_out ? g_out : (float*)(buf + 512) + n;
  cudaMemsetAsync(buf, 0, 512, s);
  cudaMemsetAsync(f, 0, (size_t)n * 4, s);
  cudaMemsetAsync(g, 0, (size_t)m * 4, s);
  double eps = epsilon;
  if (!(eps > 0)) {
    sinkhorn_mean_kernel<<<148 * 8, 256, 0, s>>>(cost, (int64_t)n * m, scal);
    double sum = 0;
    if ((rc = check_cuda(cudaMemcpyAsync(&sum, scal, 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync"))) { cudaFree(buf); return rc; }
    if ((rc = check_cuda(cudaStreamSynchronize(s), "sinkhorn mean"))) { cudaFree(buf); return rc; }
    eps = 0.05 * sum / ((double)n * m);
    if (!(eps > 0)) eps = 1e-30;
  }
  const double a_i = 1.0 / n, b_j = 1.0 / m;
  const float log_a = (float)log(a_i), log_b = (float)log(b_j);
  // OTT potentials absorb the marginals (f_ott = f + eps log a); the standard ones are kept here, same fixed point
  int it = 0, converged = 0;
  double err = INFINITY;
  while (it < max_iterations && !converged) {
    for (int k = 0; k < inner_iterations && it < max_iterations; ++k, ++it) {
      const bool last = (k == inner_iterations - 1) || (it == max_iterations - 1);
      sinkhorn_cols_kernel<<<(m + 31) / 32, 256, 0, s>>>(cost, n, m, f, g, (float)eps, log_a);
      if (last) cudaMemsetAsync(scal + 1, 0, 8, s);
      sinkhorn_rows_kernel<<<n, 256, 0, s>>>(cost, n, m, g, f, (float)eps, log_b, last ? scal + 1 : nullptr, a_i);
    }
    if ((rc = check_cuda(cudaMemcpyAsync(&err, scal + 1, 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync"))) { cudaFree(buf); return rc; }
    if ((rc = check_cuda(cudaStreamSynchronize(s), "sinkhorn iteration"))) { cudaFree(buf); return rc; }
    converged = err < threshold;
  }
  cudaMemsetAsync(scal + 4, 0, 24, s);
  sinkhorn_objective_kernel<<<n, 256, 0, s>>>(cost, n, m, f, g, (float)eps, a_i, b_j, scal + 4);
  double obj[3];
  if ((rc = check_cuda(cudaMemcpyAsync(obj, scal + 4, 24, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync"))) { cudaFree(buf); return rc; }
  if ((rc = check_cuda(cudaStreamSynchronize(s), "sinkhorn objective"))) { cudaFree(buf); return rc; }
  out_host[0] = obj[0] + obj[1] + eps * (1.0 - obj[2]);
  out_host[1] = it;
  out_host[2] = err;
  out_host[3] = converged;
  out_host[4] = eps;
  cudaFree(buf);
  return AMCMC_OK;
}
