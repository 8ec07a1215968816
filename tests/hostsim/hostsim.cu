// hostsim.cu -- TEST-ONLY debugging aid: compiles the __host__ __device__ per-chain body of the
// thread-per-chain CUDA kernel (adaptive_mcmc_b200/csrc/arwmh_small.cuh) for the HOST so its logic
// can be checked against the oracle in this GPU-less container before spending GPU time.
// It is never built by __graft_entry__.build(), never loaded by the adaptive_mcmc_b200 package and
// is not a fallback: the product path has no CPU implementation.
#include "../../adaptive_mcmc_b200/csrc/arwmh_small.cuh"
#include "../../adaptive_mcmc_b200/csrc/asss_small.cuh"

#include <vector>

using namespace amcmc;

template <typename R>
static void run_es(const double* y, const double* sigma, double cst, int64_t C, R* z, R* pe, R* macc, R* loc, R* scale,
                   R* lam, R* asc, int64_t i0, int64_t n_steps, int64_t thinning, int64_t collect_start,
                   int64_t num_warmup, double lr, double target, double eps, uint64_t seed, int64_t chain_offset,
                   const R* normals, const R* uniforms, R* out_z, R* out_pe, uint8_t* out_acc, int adapt) {
  EightSchoolsModel<R> m;
  for (int j = 0; j < 8; ++j) { m.y[j] = (R)y[j]; m.inv_sigma[j] = (R)(1.0 / sigma[j]); }
  m.cst = (R)cst;
  StateView<R> st{C, z, pe, macc, loc, scale, lam, asc};
  RunView<R> a{i0, n_steps, thinning, collect_start, num_warmup, (R)lr, (R)target, (R)eps, seed, chain_offset,
               normals, uniforms, out_z, out_pe, out_acc};
  for (int64_t c = 0; c < C; ++c) {
    if (adapt) {
      if (normals) arwmh_chain_run<EightSchoolsModel<R>, R, true, true>(m, st, a, c);
      else arwmh_chain_run<EightSchoolsModel<R>, R, true, false>(m, st, a, c);
    } else {
      if (normals) arwmh_chain_run<EightSchoolsModel<R>, R, false, true>(m, st, a, c);
      else arwmh_chain_run<EightSchoolsModel<R>, R, false, false>(m, st, a, c);
    }
  }
}

#define ARGS(R) const double* y, const double* sigma, double cst, int64_t C, R* z, R* pe, R* macc, R* loc, R* scale, \
                R* lam, R* asc, int64_t i0, int64_t n_steps, int64_t thinning, int64_t collect_start,                \
                int64_t num_warmup, double lr, double target, double eps, uint64_t seed, int64_t chain_offset,       \
                const R* normals, const R* uniforms, R* out_z, R* out_pe, uint8_t* out_acc, int adapt
#define PASS y, sigma, cst, C, z, pe, macc, loc, scale, lam, asc, i0, n_steps, thinning, collect_start, num_warmup, lr, \
             target, eps, seed, chain_offset, normals, uniforms, out_z, out_pe, out_acc, adapt

extern "C" void hostsim_es_f32(ARGS(float)) { run_es<float>(PASS); }
extern "C" void hostsim_es_f64(ARGS(double)) { run_es<double>(PASS); }

template <typename R>
static void run_es_asss(const double* y, const double* sigma, double cst, int64_t C, R* z, R* pe, R* macc, R* loc, R* scale,
                        R* lam, R* asc, int64_t i0, int64_t n_steps, int64_t thinning, int64_t collect_start,
                        int64_t num_warmup, double lr, double target, double eps, uint64_t seed, int64_t chain_offset,
                        const R* normals, const R* uniforms, R* out_z, R* out_pe, uint8_t* out_acc, int adapt) {
  EightSchoolsModel<R> m;
  for (int j = 0; j < 8; ++j) { m.y[j] = (R)y[j]; m.inv_sigma[j] = (R)(1.0 / sigma[j]); }
  m.cst = (R)cst;
  StateView<R> st{C, z, pe, macc, loc, scale, lam, asc};
  RunView<R> a{i0, n_steps, thinning, collect_start, num_warmup, (R)lr, (R)target, (R)eps, seed, chain_offset,
               normals, uniforms, out_z, out_pe, out_acc};
  for (int64_t c = 0; c < C; ++c) {
    if (normals) asss_chain_run<EightSchoolsModel<R>, R, true, true>(m, st, a, c);
    else asss_chain_run<EightSchoolsModel<R>, R, false, true>(m, st, a, c);
  }
}
extern "C" void hostsim_es_asss_f32(ARGS(float)) { run_es_asss<float>(PASS); }
extern "C" void hostsim_es_asss_f64(ARGS(double)) { run_es_asss<double>(PASS); }

// The hand-off of the balanced launch (arwmh_small_balanced_kernel) on the host: the run is cut into ranges of `seg` steps and
// the chain's registers travel through a ChainSlot between ranges, exactly as a chain group moves from worker to worker.
template <typename R>
static void run_es_split(const double* y, const double* sigma, double cst, int64_t C, R* z, R* pe, R* macc, R* loc, R* scale,
                         R* lam, R* asc, int64_t i0, int64_t n_steps, int64_t thinning, int64_t collect_start,
                         int64_t num_warmup, double lr, double target, double eps, uint64_t seed, int64_t chain_offset,
                         const R* normals, const R* uniforms, R* out_z, R* out_pe, uint8_t* out_acc, int adapt, int seg, int asss) {
  using M = EightSchoolsModel<R>;
  using Slot = ChainSlot<R, M::D>;
  M m;
  for (int j = 0; j < 8; ++j) { m.y[j] = (R)y[j]; m.inv_sigma[j] = (R)(1.0 / sigma[j]); }
  m.cst = (R)cst;
  StateView<R> st{C, z, pe, macc, loc, scale, lam, asc};
  RunView<R> a{i0, n_steps, thinning, collect_start, num_warmup, (R)lr, (R)target, (R)eps, seed, chain_offset,
               normals, uniforms, out_z, out_pe, out_acc};
  std::vector<R> slot((size_t)Slot::NREG * 32);
  for (int64_t c = 0; c < C; ++c) {
    const int lane = (int)(c % 32);
    const Philox rng(a.seed, (uint64_t)(c + a.chain_offset));
    for (int64_t t0 = 0; t0 < n_steps; t0 += seg) {
      const int64_t t1 = t0 + seg < n_steps ? t0 + seg : n_steps;
      ChainRegs<R, M::D> s;
      if (t0 == 0) load_chain(s, st, c);
      else Slot::restore(s, slot.data(), lane);
      if (asss) {
        if (normals) AsssRange<M, R, true, true>::run(s, m, a, rng, C, c, t0, t1);
        else AsssRange<M, R, true, false>::run(s, m, a, rng, C, c, t0, t1);
      } else if (adapt) {
        if (normals) ArwmhRange<M, R, true, true>::run(s, m, a, rng, C, c, t0, t1);
        else ArwmhRange<M, R, true, false>::run(s, m, a, rng, C, c, t0, t1);
      } else {
        if (normals) ArwmhRange<M, R, false, true>::run(s, m, a, rng, C, c, t0, t1);
        else ArwmhRange<M, R, false, false>::run(s, m, a, rng, C, c, t0, t1);
      }
      if (t1 == n_steps) {
        if (adapt || asss) store_chain<R, M::D, true>(s, st, c);
        else store_chain<R, M::D, false>(s, st, c);
      } else {
        Slot::save(s, slot.data(), lane);
      }
    }
  }
}
extern "C" void hostsim_es_split_f32(ARGS(float), int seg, int asss) { run_es_split<float>(PASS, seg, asss); }
extern "C" void hostsim_es_split_f64(ARGS(double), int seg, int asss) { run_es_split<double>(PASS, seg, asss); }
