// eval.cu -- sample-quality metrics of the reference on the GPU (python/utils/evaluation.py), SURVEY 8f rank 4:
// the O(n m d) all-pairs passes behind `mmd_heuristic` / `mmd2_unbiased` (:201-294), the median heuristic
// bandwidth (:283), the cost matrix of the 1-1 Wasserstein coupling (:58) and the moment estimates (:33-34).
//
// Samples are row-major [n][d] float32, as the reference's jnp arrays.  One tile loop serves all passes: a thread
// owns one row of x and keeps TJ running squared distances (or p-norm sums) in registers while the dimension is
// swept in chunks of 32 -- x chunk in registers, y chunk broadcast from shared memory -- so any d works and nothing
// of size n*m is ever stored (except the cost matrix, which is the output).  The median of the m^2 squared distances
// is an exact 3-pass radix select on the float bit patterns (11 + 11 + 10 bits): each pass recomputes the distances
// and histograms the keys that match the prefix found so far; no sort, no m^2 buffer.
#include <cmath>
#include <cstring>
#include "internal.h"

namespace amcmc {

constexpr int EV_TI = 128;  // rows of x per block (one per thread)
constexpr int EV_TJ = 32;   // rows of y per tile
constexpr int EV_KC = 32;   // dimension chunk

// acc[j] = sum_k f(x_ik - y_jk) for the tile (j0 .. j0+TJ); `post(i, j, acc)` is called for every valid pair.
template <class Term, class Post>
__device__ __forceinline__ void pairwise_tiles(const float* __restrict__ x, int64_t n, const float* __restrict__ y, int64_t m, int d,
                                               int64_t j_begin, int64_t j_end, float (*ys)[EV_TJ], Term term, Post post) {
  const int64_t i = (int64_t)blockIdx.x * EV_TI + threadIdx.x;
  const int64_t ic = i < n ? i : n - 1;
  const bool one_chunk = d <= EV_KC;  // the usual case (d = 4, 10, 26): the row of x is loaded once per block
  float xr[EV_KC];
  if (one_chunk) {
#pragma unroll
    for (int k = 0; k < EV_KC; ++k) xr[k] = (k < d) ? x[ic * d + k] : 0.f;
  }
  for (int64_t j0 = j_begin; j0 < j_end; j0 += EV_TJ) {
    float acc[EV_TJ];
#pragma unroll
    for (int j = 0; j < EV_TJ; ++j) acc[j] = 0.f;
    for (int k0 = 0; k0 < d; k0 += EV_KC) {
      if (!one_chunk) {
#pragma unroll
        for (int k = 0; k < EV_KC; ++k) xr[k] = (k0 + k < d) ? x[ic * d + k0 + k] : 0.f;
      }
      __syncthreads();
      for (int e = threadIdx.x; e < EV_KC * EV_TJ; e += EV_TI) {
        const int j = e / EV_KC, k = e % EV_KC;  // consecutive threads read consecutive k of one row: coalesced
        const int64_t jj = j0 + j;
        ys[k][j] = (jj < m && k0 + k < d) ? y[jj * d + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < EV_KC; ++k) {
#pragma unroll
        for (int j = 0; j < EV_TJ; ++j) acc[j] = term(xr[k] - ys[k][j], acc[j]);
      }
    }
    if (i < n) {
#pragma unroll
      for (int j = 0; j < EV_TJ; ++j)
        if (j0 + j < j_end) post(i, j0 + j, acc[j]);
    }
  }
}

struct SqTerm {
  __device__ __forceinline__ float operator()(float df, float a) const { return fmaf(df, df, a); }
};
struct AbsTerm {
  __device__ __forceinline__ float operator()(float df, float a) const { return a + fabsf(df); }
};
struct PowTerm {
  float p;
  __device__ __forceinline__ float operator()(float df, float a) const { return a + powf(fabsf(df), p); }
};

__device__ __forceinline__ void split_range(int64_t m, int64_t& j_begin, int64_t& j_end) {
  const int64_t tiles = (m + EV_TJ - 1) / EV_TJ;
  const int64_t per = (tiles + gridDim.y - 1) / gridDim.y;
  j_begin = (int64_t)blockIdx.y * per * EV_TJ;
  j_end = j_begin + per * EV_TJ;
  if (j_end > m) j_end = m;
  if (j_begin > m) j_begin = m;
}

// out += sum_ij exp(-gamma |x_i - y_j|^2)   (gaussian_kernel, evaluation.py:201-222; fp32 per pair, fp64 totals)
__global__ void __launch_bounds__(EV_TI) eval_kernel_sum_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ y, int64_t m,
                                                                 int d, float gamma, int skip_diag, double* __restrict__ out) {
  __shared__ float ys[EV_KC][EV_TJ];
  __shared__ double red[EV_TI / 32];
  int64_t jb, je;
  split_range(m, jb, je);
  double tot = 0.0;
  float part = 0.f;
  pairwise_tiles(x, n, y, m, d, jb, je, ys, SqTerm(), [&](int64_t i, int64_t j, float s) {
    const float e = expf(-gamma * s);
    part += (skip_diag && i == j) ? 0.f : e;
    if ((j & (EV_TJ - 1)) == EV_TJ - 1) { tot += (double)part; part = 0.f; }  // fp32 partials stay one tile long
  });
  tot += (double)part;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < EV_TI / 32; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

// histogram of ((key >> shift) & (nbins - 1)) over the pairs whose key matches `prefix` above `shift + bits`
__global__ void __launch_bounds__(EV_TI) eval_sqdist_hist_kernel(const float* __restrict__ y, int64_t m, int d, uint32_t prefix, int prefix_shift,
                                                                  int shift, int nbins, unsigned long long* __restrict__ hist) {
  __shared__ float ys[EV_KC][EV_TJ];
  __shared__ unsigned int sh[2048];
  for (int b = threadIdx.x; b < nbins; b += EV_TI) sh[b] = 0u;
  __syncthreads();
  int64_t jb, je;
  split_range(m, jb, je);
  pairwise_tiles(y, m, y, m, d, jb, je, ys, SqTerm(), [&](int64_t, int64_t, float s) {
    const uint32_t key = __float_as_uint(s);
    if (prefix_shift >= 32 || (key >> prefix_shift) == prefix) atomicAdd(&sh[(key >> shift) & (uint32_t)(nbins - 1)], 1u);
  });
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += EV_TI)
    if (sh[b]) atomicAdd(&hist[b], (unsigned long long)sh[b]);
}

// out[i][j] = |x_i - y_j|_p   (scipy.spatial.distance_matrix(u, v, p=ord), evaluation.py:58)
template <class Term>
__global__ void __launch_bounds__(EV_TI) eval_cost_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ y, int64_t m, int d,
                                                           Term term, float inv_p, int mode, float* __restrict__ out) {
  __shared__ float ys[EV_KC][EV_TJ];
  int64_t jb, je;
  split_range(m, jb, je);
  pairwise_tiles(x, n, y, m, d, jb, je, ys, term, [&](int64_t i, int64_t j, float s) {
    out[i * m + j] = mode == 2 ? sqrtf(s) : (mode == 1 ? s : powf(s, inv_p));
  });
}

// out[k] += sum_i x_ik^p   (pth_moment_rmse, evaluation.py:33-34)
__global__ void eval_moment_kernel(const float* __restrict__ x, int64_t n, int d, float p, double* __restrict__ out) {
  const int k = blockIdx.y;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i * d + k];
    s += (double)(p == 2.f ? v * v : (p == 1.f ? v : powf(v, p)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&out[k], s);
}

static dim3 pair_grid(int64_t n, int64_t m) {
  const int64_t gx = (n + EV_TI - 1) / EV_TI;
  const int64_t tiles = (m + EV_TJ - 1) / EV_TJ;
  int64_t gy = (148 * 8 + gx - 1) / gx;  // enough blocks for a few waves on 148 SMs
  if (gy > tiles) gy = tiles;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  return dim3((unsigned)gx, (unsigned)gy);
}

// persistent per-device scratch (2048 histogram bins + up to 2048 doubles): cudaMalloc / cudaFree per call would
// cost more than the kernels.  Calls on one device are expected from one host thread at a time (as in the reference).
static void* eval_scratch() {
  static void* buf[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!buf[dev] && cudaMalloc(&buf[dev], 2048 * sizeof(unsigned long long)) != cudaSuccess) buf[dev] = nullptr;
  return buf[dev];
}

static int eval_args_ok(const void* x, int64_t n, const void* y, int64_t m, int d) {
  if (!x || !y || n <= 0 || m <= 0 || d <= 0) { set_error("evaluation: null or empty sample array"); return 0; }
  if ((n + EV_TI - 1) / EV_TI > 0x7fffffff) { set_error("evaluation: too many samples"); return 0; }
  return 1;
}

}  // namespace amcmc

using namespace amcmc;

extern "C" int amcmc_eval_kernel_sum(const float* x, int64_t n, const float* y, int64_t m, int d, double gamma, int skip_diagonal,
                                     double* out_host, void* stream) {
  if (!eval_args_ok(x, n, y, m, d) || !out_host) { if (!out_host) set_error("evaluation: out_host is null"); return AMCMC_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = (double*)eval_scratch();
  int rc;
  if (!acc) { set_error("evaluation: scratch allocation failed"); return AMCMC_ERR_CUDA; }
  cudaMemsetAsync(acc, 0, sizeof(double), s);
  eval_kernel_sum_kernel<<<pair_grid(n, m), EV_TI, 0, s>>>(x, n, y, m, d, (float)gamma, skip_diagonal, acc);
  rc = check_cuda(cudaGetLastError(), "eval_kernel_sum_kernel launch");
  if (!rc) rc = check_cuda(cudaMemcpyAsync(out_host, acc, sizeof(double), cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync");
  if (!rc) rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  return rc;
}

extern "C" int amcmc_eval_sqdist_median(const float* y, int64_t m, int d, double* out_host, void* stream) {
  if (!eval_args_ok(y, m, y, m, d) || !out_host) { if (!out_host) set_error("evaluation: out_host is null"); return AMCMC_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* hist = (unsigned long long*)eval_scratch();
  int rc = 0;
  if (!hist) { set_error("evaluation: scratch allocation failed"); return AMCMC_ERR_CUDA; }
  unsigned long long hh[2048];
  const unsigned long long total = (unsigned long long)m * (unsigned long long)m;
  // jnp.median of an even count = mean of the two middle order statistics
  const unsigned long long ranks[2] = {(total - 1) / 2, total / 2};
  float vals[2] = {0.f, 0.f};
  // The upper middle rank is the next order statistic: it is found in the histograms of the lower one as long as it stays in the
  // same bin (or, in the last pass, in the next non-empty bin of the same prefix); only when the two ranks part in an earlier
  // pass (their keys differ above bit 10) is the select run again for it.
  bool second_known = ranks[1] == ranks[0];
  for (int r = 0; r < 2 && !rc; ++r) {
    if (r == 1 && second_known) { if (ranks[1] == ranks[0]) vals[1] = vals[0]; break; }
    unsigned long long k = ranks[r];
    uint32_t prefix = 0;
    bool together = (r == 0) && !second_known;  // rank k + 1 still shares the prefix
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    for (int pass = 0; pass < 3 && !rc; ++pass) {
      const int nb = 1 << bits[pass];
      cudaMemsetAsync(hist, 0, nb * sizeof(unsigned long long), s);
      eval_sqdist_hist_kernel<<<pair_grid(m, m), EV_TI, 0, s>>>(y, m, d, prefix, shifts[pass] + bits[pass], shifts[pass], nb, hist);
      if ((rc = check_cuda(cudaGetLastError(), "eval_sqdist_hist_kernel launch"))) break;
      if ((rc = check_cuda(cudaMemcpyAsync(hh, hist, nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync"))) break;
      if ((rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize"))) break;
      int b = 0;
      unsigned long long cum = 0;
      for (; b < nb; ++b) {
        if (cum + hh[b] > k) break;
        cum += hh[b];
      }
      if (b == nb) { set_error("evaluation: median select ran past the histogram (NaN distances?)"); rc = AMCMC_ERR_ARG; break; }
      if (together && k + 1 >= cum + hh[b]) {  // rank k + 1 is the first key of a later bin
        together = false;
        if (pass == 2) {                       // all 32 bits resolved: the next non-empty bin IS the value
          int b2 = b + 1;
          while (b2 < nb && hh[b2] == 0) ++b2;
          if (b2 < nb) {
            const uint32_t key = (prefix << bits[pass]) | (uint32_t)b2;
            std::memcpy(&vals[1], &key, 4);
            second_known = true;
          }
        }
      }
      k -= cum;
      prefix = (prefix << bits[pass]) | (uint32_t)b;
    }
    std::memcpy(&vals[r], &prefix, 4);
    if (r == 0 && together) { vals[1] = vals[0]; second_known = true; }
  }
  if (!rc) *out_host = (double)(vals[0] + (vals[1] - vals[0]) * 0.5f);
  return rc;
}

extern "C" int amcmc_eval_cost_matrix(const float* x, int64_t n, const float* y, int64_t m, int d, double ord, float* out, void* stream) {
  if (!eval_args_ok(x, n, y, m, d) || !out || !(ord >= 1.0)) {
    if (!out) set_error("evaluation: out is null");
    else if (!(ord >= 1.0)) set_error("evaluation: norm order must be >= 1");
    return AMCMC_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (ord == 2.0) eval_cost_kernel<<<pair_grid(n, m), EV_TI, 0, s>>>(x, n, y, m, d, SqTerm(), 0.5f, 2, out);
  else if (ord == 1.0) eval_cost_kernel<<<pair_grid(n, m), EV_TI, 0, s>>>(x, n, y, m, d, AbsTerm(), 1.f, 1, out);
  else eval_cost_kernel<<<pair_grid(n, m), EV_TI, 0, s>>>(x, n, y, m, d, PowTerm{(float)ord}, (float)(1.0 / ord), 0, out);
  return check_cuda(cudaGetLastError(), "eval_cost_kernel launch");
}

extern "C" int amcmc_eval_moment(const float* x, int64_t n, int d, double p, double* out_host, void* stream) {
  if (!x || n <= 0 || d <= 0 || d > 2048 || !out_host) { set_error("evaluation: bad arguments to amcmc_eval_moment"); return AMCMC_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = (double*)eval_scratch();
  int rc;
  if (!acc) { set_error("evaluation: scratch allocation failed"); return AMCMC_ERR_CUDA; }
  cudaMemsetAsync(acc, 0, sizeof(double) * d, s);
  int gx = (int)((n + 255) / 256);
  if (gx > 592) gx = 592;
  eval_moment_kernel<<<dim3(gx, d), 256, 0, s>>>(x, n, d, (float)p, acc);
  rc = check_cuda(cudaGetLastError(), "eval_moment_kernel launch");
  if (!rc) rc = check_cuda(cudaMemcpyAsync(out_host, acc, sizeof(double) * d, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync");
  if (!rc) rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  if (!rc)
    for (int k = 0; k < d; ++k) out_host[k] /= (double)n;
  return rc;
}
