// common.cuh -- scalar-type traits, math wrappers and the Philox4x32-10 RNG shared by all kernels.
// Everything here is __host__ __device__ so tests can also compile the step logic for the host
// (tests/hostsim, debugging aid only -- never a product path).
#pragma once
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

#define AMCMC_HD __host__ __device__ __forceinline__

namespace amcmc {

// ---------------------------------------------------------------------------------------------
// Math wrappers.  fp32 on the device uses the SFU (MUFU) approximations: 1-2 ulp, which is far
// inside the 1e-3 fp32 trajectory tolerance the north star states, and keeps the fused step
// issue-bound instead of bound by IEEE div/sqrt fix-up sequences.  fp64 is the exact parity path.
// ---------------------------------------------------------------------------------------------
template <typename R> struct Num;

template <> struct Num<float> {
  static AMCMC_HD float inf() { return INFINITY; }
  static AMCMC_HD float exp(float x) {
#ifdef __CUDA_ARCH__
    return __expf(x);
#else
    return ::expf(x);
#endif
  }
  static AMCMC_HD float log(float x) {
#ifdef __CUDA_ARCH__
    return __logf(x);
#else
    return ::logf(x);
#endif
  }
  // ln(x) for x known to be a NORMAL positive float (no denormal fix-up sequence): one MUFU.LG2 + one FMUL
  static AMCMC_HD float log_normal_range(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.6931471805599453f;
#else
    return ::logf(x);
#endif
  }
  static AMCMC_HD float log1p(float x) { return ::log1pf(x); }
  static AMCMC_HD float rcp(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
  }
  // sqrt for x >= 0 (returns 0 at 0)
  static AMCMC_HD float sqrt(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return ::sqrtf(x);
#endif
  }
  // n^(-p) for n >= 1
  static AMCMC_HD float pow_neg(float n, float p) {
#ifdef __CUDA_ARCH__
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(n));   // n >= 1: normal range, lg2(1) == 0 exactly
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-p * l));
    return r;
#else
    return 1.0f / ::powf(n, p);
#endif
  }
  static AMCMC_HD bool isnan(float x) { return x != x; }
  static AMCMC_HD float abs(float x) { return ::fabsf(x); }
  static constexpr float kBig = 1e15f;  // |delta| guard for the rank-1 sweep (delta^2 must not overflow)
};

template <> struct Num<double> {
  static AMCMC_HD double inf() { return (double)INFINITY; }
  static AMCMC_HD double exp(double x) { return ::exp(x); }
  static AMCMC_HD double log(double x) { return ::log(x); }
  static AMCMC_HD double log1p(double x) { return ::log1p(x); }
  static AMCMC_HD double rcp(double x) { return 1.0 / x; }
  static AMCMC_HD double sqrt(double x) { return ::sqrt(x); }
  static AMCMC_HD double pow_neg(double n, double p) { return 1.0 / ::pow(n, p); }
  static AMCMC_HD bool isnan(double x) { return x != x; }
  static AMCMC_HD double abs(double x) { return ::fabs(x); }
  static constexpr double kBig = 1e150;
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Stream layout (mirrored by oracle/arwmh_numpy.py
// philox_words and oracle/arwmh_oracle.c):
//   key     = (seed_lo, seed_hi)
//   counter = (step_lo, (step_hi & 0xFFFFFF) | (blk << 24), chain_lo, chain_hi)
// `step` is the global iteration index ARWMHState.i, `chain` the global chain id, so a run is
// reproducible independently of how it is cut into launches or sharded over GPUs.
// ---------------------------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1;      // seed
  uint32_t ch0, ch1;    // chain id
  AMCMC_HD Philox(uint64_t seed, uint64_t chain)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), ch0((uint32_t)chain), ch1((uint32_t)(chain >> 32)) {}

  static AMCMC_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
  }

  AMCMC_HD void block(uint64_t step, uint32_t blk, uint32_t (&out)[4]) const {
    uint32_t c0 = (uint32_t)step;
    uint32_t c1 = ((uint32_t)(step >> 32) & 0xFFFFFFu) | (blk << 24);
    uint32_t c2 = ch0, c3 = ch1;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ a;
      uint32_t n2 = hi0 ^ c3 ^ b;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// Box-Muller on one word pair, in float32 (also for the fp64 kernels: the draws are the same
// numbers in both precisions).  u1 in (0,1], theta in [-pi, pi).
AMCMC_HD void box_muller(uint32_t wa, uint32_t wb, float& z0, float& z1) {
  float u1 = fmaf((float)wa, 0x1p-32f, 0x1p-33f);
  float th = 6.283185307179586f * ((float)(int32_t)wb * 0x1p-32f);
#ifdef __CUDA_ARCH__
  float r = Num<float>::sqrt(-2.0f * Num<float>::log_normal_range(u1));  // u1 >= 2^-33: never denormal
  float s, c;
  __sincosf(th, &s, &c);
#else
  float r = ::sqrtf(-2.0f * ::logf(u1));
  float s = ::sinf(th), c = ::cosf(th);
#endif
  z0 = r * c;
  z1 = r * s;
}

AMCMC_HD float word_to_uniform(uint32_t w) { return (float)(w >> 8) * 0x1p-24f; }

// Draws for one step of a D-dimensional chain: z[0..D) ~ N(0,1), u ~ U[0,1).
// Word usage: pairs (2k, 2k+1) -> (z_2k, z_2k+1); word 2*ceil(D/2) -> u.
template <typename R, int D>
AMCMC_HD void philox_draws(const Philox& g, uint64_t step, R (&z)[D], R& u) {
  constexpr int NPAIR = (D + 1) / 2;
  constexpr int NW = 2 * NPAIR + 1;
  constexpr int NBLK = (NW + 3) / 4;
  uint32_t w[NBLK * 4];
#pragma unroll
  for (int b = 0; b < NBLK; ++b) {
    uint32_t o[4];
    g.block(step, (uint32_t)b, o);
    w[4 * b] = o[0]; w[4 * b + 1] = o[1]; w[4 * b + 2] = o[2]; w[4 * b + 3] = o[3];
  }
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) {
    float a, b;
    box_muller(w[2 * p], w[2 * p + 1], a, b);
    z[2 * p] = (R)a;
    if (2 * p + 1 < D) z[2 * p + 1] = (R)b;
  }
  u = (R)word_to_uniform(w[2 * NPAIR]);
}

constexpr uint64_t kInitStep = ((uint64_t)1 << 56) - 1;  // reserved iteration index for q0 draws

}  // namespace amcmc
