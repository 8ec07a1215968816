// jax_rng.cu -- the reference's random stream on the GPU.
//
// The reference draws from jax.random with the default threefry2x32 PRNG (python/kernels/arwmh.py:162-165,174):
//     rng_key, key_proposal, key_accept = random.split(rng_key, 3)
//     prop_base = dist.Normal().sample(key_proposal, sample_shape=(dim,))      # random.normal(key, (dim,))
//     u         = dist.Uniform().sample(key_accept)                            # random.uniform(key, ())
// This kernel generates exactly that stream for C chains x T steps -- one thread per chain, the chain's key carried from
// step to step -- into the external-draws layout of amcmc_arwmh_run (normals[T][d][C], uniforms[T][C]), so that a run started
// from a JAX key consumes the numbers the reference would (uniforms bit for bit; normals to the last bits of log1p).
// Every kernel family (register, shared-memory, tensor-core) can then replay a reference run through rng_mode EXTERNAL.
//
// Restated from the published algorithms (JAX is a dependency of the reference, not part of it; not installable here):
// Threefry-2x32-20 (Salmon et al., SC'11); jax._src.prng `threefry_2x32` counter pairing (first half / second half of the
// counter array, odd lengths padded with one zero), `_threefry_split_original`, `_threefry_random_bits_original` (the
// non-partitionable default before jax 0.5); jax._src.random `uniform` (23 mantissa bits) and `normal`
// (sqrt(2) * erfinv(U(-1 + ulp, 1))) with XLA's float32 erfinv polynomial (Giles 2012).  Test: tests/test_gpu_jax_rng.py
// against oracle/jax_random.py, which is pinned to the Random123 vectors and to the values the JAX documentation prints.
#include <cmath>
#include <cstdint>
#include "internal.h"

namespace amcmc {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int blk = 0; blk < 5; ++blk) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = (blk & 1) ? (q == 0 ? 17 : q == 1 ? 29 : q == 2 ? 16 : 24) : (q == 0 ? 13 : q == 1 ? 15 : q == 2 ? 26 : 6);
      x0 += x1;
      x1 = rotl32(x1, r);
      x1 ^= x0;
    }
    x0 += ks[(blk + 1) % 3];
    x1 += ks[(blk + 2) % 3] + (uint32_t)(blk + 1);
  }
}

// element `idx` of jax's threefry_2x32(key, iota(n)): the counter array (padded to even length with a zero) is cut in two
// halves; block i hashes (counts[i], counts[h + i]) and its two outputs are elements i and h + i
__device__ __forceinline__ uint32_t jax_bits(uint32_t k0, uint32_t k1, int idx, int n) {
  const int h = (n + 1) >> 1;
  const int i = idx < h ? idx : idx - h;
  uint32_t x0 = (uint32_t)i;
  uint32_t x1 = (h + i < n) ? (uint32_t)(h + i) : 0u;  // the pad of an odd-length array is the counter value 0
  threefry2x32(k0, k1, x0, x1);
  return idx < h ? x0 : x1;
}

__device__ __forceinline__ float jax_bits_to_unit(uint32_t bits) { return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f; }

// a * b + c with two roundings (no FMA contraction): what an unfused float32 evaluation -- NumPy, XLA-CPU without fast math -- gives
__device__ __forceinline__ float mul_add(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }

__device__ __forceinline__ float xla_erfinv_f32(float x) {
  float w = -log1pf(-__fmul_rn(x, x));
  float p;
  if (w < 5.0f) {
    w -= 2.5f;
    p = 2.81022636e-08f;
    p = mul_add(p, w, 3.43273939e-07f);
    p = mul_add(p, w, -3.5233877e-06f);
    p = mul_add(p, w, -4.39150654e-06f);
    p = mul_add(p, w, 0.00021858087f);
    p = mul_add(p, w, -0.00125372503f);
    p = mul_add(p, w, -0.00417768164f);
    p = mul_add(p, w, 0.246640727f);
    p = mul_add(p, w, 1.50140941f);
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = mul_add(p, w, 0.000100950558f);
    p = mul_add(p, w, 0.00134934322f);
    p = mul_add(p, w, -0.00367342844f);
    p = mul_add(p, w, 0.00573950773f);
    p = mul_add(p, w, -0.0076224613f);
    p = mul_add(p, w, 0.00943887047f);
    p = mul_add(p, w, 1.00167406f);
    p = mul_add(p, w, 2.83297682f);
  }
  return fabsf(x) == 1.0f ? copysignf(INFINITY, x) : __fmul_rn(p, x);
}

template <typename R>
__global__ void jax_draws_kernel(uint32_t* __restrict__ keys, int64_t C, int d, int64_t T, R* __restrict__ normals,
                                 R* __restrict__ uniforms) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  uint32_t k0 = keys[c], k1 = keys[C + c];
  const float lo = -0.99999994f;           // nextafter(-1, 0)
  const float span = 1.0f - lo;            // rounds to 2.0f in float32, as in jax
  for (int64_t t = 0; t < T; ++t) {
    // split(key, 3): threefry_2x32(key, iota(6)) reshaped (3, 2): blocks (0,3), (1,4), (2,5) -> y0[0..2], y1[0..2];
    // key' = (y0[0], y0[1]), key_proposal = (y0[2], y1[0]), key_accept = (y1[1], y1[2])
    uint32_t a0 = 0, b0 = 3, a1 = 1, b1 = 4, a2 = 2, b2 = 5;
    threefry2x32(k0, k1, a0, b0);
    threefry2x32(k0, k1, a1, b1);
    threefry2x32(k0, k1, a2, b2);
    const uint32_t kp0 = a2, kp1 = b0, ka0 = b1, ka1 = b2;
    k0 = a0;
    k1 = a1;
    for (int k = 0; k < d; ++k) {
      const float u = fmaxf(lo, mul_add(jax_bits_to_unit(jax_bits(kp0, kp1, k, d)), span, lo));
      normals[(t * d + k) * C + c] = (R)__fmul_rn(1.41421354f, xla_erfinv_f32(u));
    }
    uniforms[t * C + c] = (R)fmaxf(0.0f, jax_bits_to_unit(jax_bits(ka0, ka1, 0, 1)));
  }
  keys[c] = k0;
  keys[C + c] = k1;
}

}  // namespace amcmc

using namespace amcmc;

extern "C" int amcmc_jax_draws(uint32_t* keys, int64_t n_chains, int dim, int64_t n_steps, int dtype, void* normals, void* uniforms,
                               void* stream) {
  if (!keys || !normals || !uniforms || n_chains < 1 || dim < 1 || n_steps < 0 || (dtype != AMCMC_F32 && dtype != AMCMC_F64)) {
    set_error("amcmc_jax_draws: bad argument");
    return AMCMC_ERR_ARG;
  }
  if (n_steps == 0) return AMCMC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n_chains + 127) / 128);
  if (dtype == AMCMC_F32) jax_draws_kernel<float><<<grid, 128, 0, s>>>(keys, n_chains, dim, n_steps, (float*)normals, (float*)uniforms);
  else jax_draws_kernel<double><<<grid, 128, 0, s>>>(keys, n_chains, dim, n_steps, (double*)normals, (double*)uniforms);
  return check_cuda(cudaGetLastError(), "jax_draws_kernel launch");
}
