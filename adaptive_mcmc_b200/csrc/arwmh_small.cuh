// arwmh_small.cuh -- thread-per-chain fused ARWMH kernel for small dimensions (d <= ~12).
//
// One thread owns one chain for the whole launch: position, energy, running mean, the proposal
// factor and the step size live in REGISTERS across all K fused steps; HBM is touched once at
// launch start / end (struct-of-arrays, chain index fastest => every access is coalesced) plus
// the thinned sample stream.  This is what moves eight_schools off the 624 B/step HBM roofline
//  onto the FP32/SFU issue rate.
//
// The proposal factor is carried as LDL^T:  L = Lt * diag(sqrt(Dg)),  Lt unit lower triangular.
// The reference (python/kernels/arwmh.py:190 -> NumPyro cholesky_update) converts L -> (L/diag, diag^2),
// runs the rank-one recurrence and converts back on EVERY step; carrying (Lt, Dg) across the
// fused steps removes both conversions and all but one reciprocal per column.  For K = 1 the
// arithmetic is the reference's sequence exactly.
#pragma once
#include "common.cuh"
#include "models.cuh"


namespace amcmc {

template <typename R, int D> struct ChainRegs {
  static constexpr int NL = D * (D - 1) / 2;
  R x[D];
  R mu[D];
  R Lt[NL > 0 ? NL : 1];  // strictly-lower part of the unit-diagonal factor, row-major packed
  R Dg[D];                // squared diagonal
  R U, lam, macc, asc;
};

AMCMC_HD constexpr int tri_strict(int i, int j) { return i * (i - 1) / 2 + j; }  // i > j
AMCMC_HD constexpr int tri_full(int i, int j) { return i * (i + 1) / 2 + j; }    // i >= j

// Rank-one update  Lt Dg Lt^T <- (1-gamma) Lt Dg Lt^T + gamma w w^T   (Gill-Golub-Murray-Saunders
// C1 recurrence, algebraically NumPyro's cholesky_update(sqrt(1-gamma) L, w, gamma) with t = 1/b):
//   g = D_j + c w_j^2 t;  D_j' = g;  coef = c w_j t / g;  t <- D_j t / g;
//   w_i -= w_j Lt_ij;  Lt_ij += coef w_i   (i > j)
// WANT additionally accumulates |L' e^lam' - L e^lam|_F^2 (arwmh.py:197) on the fly.
// (The additive form of the recurrence, b_{j+1} = b_j + gamma p_j^2 / D_j with every reciprocal off the column-to-column
// chain, was measured in round 2: 10 more MUFU per step, 3.27e10 instead of 3.29e10 chain-steps/s, and no change of the
// single-chain latency (0.94 us per step either way) -- not kept.)
//
// `bump` (0 or 1) is added to the pivot before its reciprocal.  The hot loop runs the sweep UNCONDITIONALLY: when the
// reference would keep the old factor (arwmh.py:191) the caller passes gamma = 0, w = 0, bump = 1, and every expression
// below then reproduces the old factor bit for bit (Dj = D, g = D, coef = 0 * finite = 0, Ln = Lo) -- no branch, no
// second register copy of the factor.  The bump only keeps 0 * rcp(0) = NaN out when a pivot is exactly zero.
template <typename R, int D, bool WANT>
AMCMC_HD R rank1_sweep(ChainRegs<R, D>& s, R (&w)[D], R gamma, R el_old, R el_new, R bump = (R)0) {
  R t = (R)1;
  const R omg = (R)1 - gamma;
  R ss = (R)0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const R Dold = s.Dg[j];
    const R Dj = omg * Dold;
    const R wj = w[j];
    const R cw = gamma * wj;
    const R g = fma(cw * wj, t, Dj);
    const R tr = t * Num<R>::rcp(g + bump);
    const R coef = cw * tr;
    t = Dj * tr;
    s.Dg[j] = g;
    R so = 0, sn = 0;
    if (WANT) {
      so = Num<R>::sqrt(Dold) * el_old;
      sn = Num<R>::sqrt(g) * el_new;
      const R dd = sn - so;
      ss = fma(dd, dd, ss);
    }
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      const R Lo = s.Lt[tri_strict(i, j)];
      w[i] = fma(-wj, Lo, w[i]);
      const R Ln = fma(coef, w[i], Lo);
      s.Lt[tri_strict(i, j)] = Ln;
      if (WANT) {
        const R df = fma(Ln, sn, -(Lo * so));
        ss = fma(df, df, ss);
      }
    }
  }
  return ss;
}

// |L|_F^2 of the current factor (used when the factor is kept: as_change = |e^lam' - e^lam| |L|_F)
template <typename R, int D> AMCMC_HD R factor_frob2(const ChainRegs<R, D>& s) {
  R ss = 0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    R col = (R)1;
#pragma unroll
    for (int i = j + 1; i < D; ++i) col = fma(s.Lt[tri_strict(i, j)], s.Lt[tri_strict(i, j)], col);
    ss = fma(s.Dg[j], col, ss);
  }
  return ss;
}

// One ARWMH.sample (python/kernels/arwmh.py:140-207) for the chain held in `s`.
// WANT_ASC: also compute as_change (:197) -- only the last step of a launch does (state snapshots are launch boundaries),
// so the hot loop instantiates WANT_ASC = false, which is straight-line code.
template <class Model, typename R, bool ADAPT, bool WANT_ASC>
AMCMC_HD bool arwmh_step(ChainRegs<R, Model::D>& s, const Model& m, const R (&z)[Model::D], R u, R nf,
                         bool n_is_one, R lr_decay, R target, R eps) {
  constexpr int D = Model::D;
  const R el = Num<R>::exp(s.lam);
  // :166-167  x' = x + (L e^lam + eps I) z,  L z = Lt (sqrt(Dg) .* z)
  R y[D], xp[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    y[j] = z[j] * Num<R>::sqrt(s.Dg[j]);
    xp[j] = fma(eps, z[j], s.x[j]);  // the eps I z term is folded in here so that z dies before the matvec
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    R acc = y[i];
#pragma unroll
    for (int j = 0; j < i; ++j) acc = fma(s.Lt[tri_strict(i, j)], y[j], acc);
    xp[i] = fma(el, acc, xp[i]);
  }
  // :170-171
  R Up = m.potential(xp);
  if (Num<R>::isnan(Up)) Up = Num<R>::inf();
  // :173-178   clip(exp(.), max=1) keeps NaN
  const R e = Num<R>::exp(s.U - Up);
  const R alpha = (e > (R)1) ? (R)1 : e;
  const bool acc = u < alpha;
#pragma unroll
  for (int k = 0; k < D; ++k) s.x[k] = acc ? xp[k] : s.x[k];
  s.U = acc ? Up : s.U;
  // :185
  s.macc = fma(alpha - s.macc, Num<R>::rcp(nf), s.macc);
  if (!ADAPT) return acc;
  // :183, :188-193
  const R gamma = n_is_one ? (R)1 : Num<R>::pow_neg(nf, lr_decay);
  R w[D];
  // pre-condition of the sweep in two comparisons: sum |delta| finite and below kBig, min D > 0
  R dabs = 0, dmin = s.Dg[0];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const R dl = s.x[k] - s.mu[k];
    s.mu[k] = fma(gamma, dl, s.mu[k]);
    w[k] = dl;
    dabs += Num<R>::abs(dl);
    dmin = s.Dg[k] < dmin ? s.Dg[k] : dmin;
  }
  const bool ok = !n_is_one && (dabs < Num<R>::kBig) && (dmin > (R)0);
  const R lam_new = fma(gamma, alpha - target, s.lam);
  // :190-191 rank-one update; "NaN => keep the old factor" is applied as a pre-condition
  // (gamma == 1 <=> zero scaled diagonal, non-positive pivot, non-finite delta), see DESIGN.md.
  if (WANT_ASC) {
    const R el_new = Num<R>::exp(lam_new);
    R ss;
    if (ok) {
      ss = rank1_sweep<R, D, true>(s, w, gamma, el, el_new);
    } else {
      const R de = el_new - el;
      ss = de * de * factor_frob2(s);
    }
    s.asc = Num<R>::sqrt(ss);  // :197
  } else {
    // branch-free: a kept factor is the sweep with gamma = 0, w = 0 (see rank1_sweep)
#pragma unroll
    for (int k = 0; k < D; ++k) w[k] = ok ? w[k] : (R)0;
    rank1_sweep<R, D, false>(s, w, ok ? gamma : (R)0, el, el, ok ? (R)0 : (R)1);
  }
  s.lam = lam_new;
  return acc;
}

// ---------------------------------------------------------------------------------------------
// Device-side views of amcmc_state / amcmc_run_args
// ---------------------------------------------------------------------------------------------
template <typename R> struct StateView {
  int64_t C;
  R *z, *pe, *macc, *loc, *scale, *lam, *asc;
};

template <typename R> struct RunView {
  int64_t i0, n_steps, thinning, collect_start, num_warmup;
  R lr_decay, target, eps;
  uint64_t seed;
  int64_t chain_offset;
  const R* normals;
  const R* uniforms;
  R* out_z;
  R* out_pe;
  uint8_t* out_acc;
  const R* ring = nullptr;  // arwmh_small_duo_kernel: shared-memory ring [2][D + 1][32] filled by the CTA's producer warp
};

template <typename R, int D>
AMCMC_HD void load_chain(ChainRegs<R, D>& s, const StateView<R>& st, int64_t c) {
  const int64_t C = st.C;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    s.x[k] = st.z[k * C + c];
    s.mu[k] = st.loc[k * C + c];
  }
  R inv_diag[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const R dg = st.scale[tri_full(j, j) * C + c];
    s.Dg[j] = dg * dg;
    inv_diag[j] = (R)1 / dg;
  }
#pragma unroll
  for (int i = 1; i < D; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) s.Lt[tri_strict(i, j)] = st.scale[tri_full(i, j) * C + c] * inv_diag[j];
  s.U = st.pe[c];
  s.lam = st.lam[c];
  s.macc = st.macc[c];
  s.asc = st.asc[c];
}

template <typename R, int D, bool ADAPT>
AMCMC_HD void store_chain(const ChainRegs<R, D>& s, const StateView<R>& st, int64_t c) {
  const int64_t C = st.C;
#pragma unroll
  for (int k = 0; k < D; ++k) st.z[k * C + c] = s.x[k];
  st.pe[c] = s.U;
  st.macc[c] = s.macc;  // frozen kernel: mean acceptance probability over this launch (pooled windows read it)
  if (!ADAPT) return;
#pragma unroll
  for (int k = 0; k < D; ++k) st.loc[k * C + c] = s.mu[k];
  R sd[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    sd[j] = ::sqrt(s.Dg[j]);
    st.scale[tri_full(j, j) * C + c] = sd[j];
  }
#pragma unroll
  for (int i = 1; i < D; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) st.scale[tri_full(i, j) * C + c] = s.Lt[tri_strict(i, j)] * sd[j];
  st.lam[c] = s.lam;
  st.asc[c] = s.asc;
}

// Draws of one step: external arrays (shared-draw parity mode) or the Philox stream.
template <typename R, int D, bool EXTERNAL, bool RING = false>
AMCMC_HD void step_draws(const RunView<R>& a, const Philox& rng, int64_t C, int64_t c, int64_t t, R (&z)[D], R& u) {
  if (RING) {
#ifdef __CUDA_ARCH__
    __syncthreads();  // the producer warp has written step t (and this warp is done with the buffer of step t - 1)
    const R* buf = a.ring + (size_t)(t & 1) * (D + 1) * 32 + (threadIdx.x & 31);
#pragma unroll
    for (int k = 0; k < D; ++k) z[k] = buf[k * 32];
    u = buf[D * 32];
#endif
  } else if (EXTERNAL) {
#pragma unroll
    for (int k = 0; k < D; ++k) z[k] = a.normals[(t * D + k) * C + c];
    u = a.uniforms[t * C + c];
  } else {
    philox_draws<R, D>(rng, (uint64_t)(a.i0 + t), z, u);
  }
}

// Steps [t0, t1) of the launch without any collection logic: the hot loop.  (Generating the draws of step t + 1 next to
// the sweep of step t -- software pipelining of the counter RNG -- was measured: 3.33e10 vs 3.36e10 chain-steps/s, and no change
// of the few-chain latency either, 0.93 us per step for 1 to 100 chains: the step's own dependent chain is the critical path;
// not kept.)  n follows arwmh.py:180-181 (it restarts at
// 1 after the warm-up); the frozen kernel (sample_Pnx, pooled windows) averages its acceptance rate over THIS launch.
template <class Model, typename R, bool ADAPT, bool EXTERNAL, bool RING = false>
AMCMC_HD void arwmh_steps(ChainRegs<R, Model::D>& s, const Model& m, const RunView<R>& a, const Philox& rng, int64_t C,
                          int64_t c, int64_t t0, int64_t t1) {
  constexpr int D = Model::D;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t i = a.i0 + t;
    R z[D], u;
    step_draws<R, D, EXTERNAL, RING>(a, rng, C, c, t, z, u);
    const int64_t n = (i < a.num_warmup) ? (i + 1) : (i + 1 - a.num_warmup);
    const R nf = ADAPT ? (R)n : (R)(t + 1);
    const bool acc = arwmh_step<Model, R, ADAPT, false>(s, m, z, u, nf, n == 1, a.lr_decay, a.target, a.eps);
    if (a.out_acc) a.out_acc[t * C + c] = (uint8_t)acc;
  }
}

// Steps [t, t_end) of the launch for one chain whose state is in `s` (host-compilable for tests/hostsim).  The range is cut
// at the collection points (numpyro.util.fori_collect as used at python/utils/kernel_utils.py:29-32: sample k = state after
// collect_start + (k+1) thinning steps) so that the hot loop carries no collection bookkeeping, and the last step of the
// LAUNCH is peeled: it alone computes as_change (arwmh.py:197).  Any split of [0, n_steps) into consecutive ranges gives
// the same trajectory and the same samples.
template <class Model, typename R, bool ADAPT, bool EXTERNAL, bool RING = false>
AMCMC_HD void arwmh_chain_range(ChainRegs<R, Model::D>& s, const Model& m, const RunView<R>& a, const Philox& rng, int64_t C,
                                int64_t c, int64_t t, int64_t t_end) {
  constexpr int D = Model::D;
  const int64_t T = a.n_steps;
  int64_t sidx = t > a.collect_start ? (t - a.collect_start) / a.thinning : 0;   // samples taken before this range
  int64_t next_collect = a.collect_start + (sidx + 1) * a.thinning;  // number of completed steps at which the next sample is taken
  while (t < t_end) {
    const int64_t seg_end = next_collect < t_end ? next_collect : t_end;
    const int64_t hot_end = seg_end < T ? seg_end : T - 1;
    arwmh_steps<Model, R, ADAPT, EXTERNAL, RING>(s, m, a, rng, C, c, t, hot_end);
    t = hot_end;
    if (seg_end == T) {  // the last step of the launch
      const int64_t i = a.i0 + t;
      R z[D], u;
      step_draws<R, D, EXTERNAL, RING>(a, rng, C, c, t, z, u);
      const int64_t n = (i < a.num_warmup) ? (i + 1) : (i + 1 - a.num_warmup);
      const R nf = ADAPT ? (R)n : (R)(t + 1);
      const bool acc = arwmh_step<Model, R, ADAPT, true>(s, m, z, u, nf, n == 1, a.lr_decay, a.target, a.eps);
      if (a.out_acc) a.out_acc[t * C + c] = (uint8_t)acc;
      t = T;
    }
    if (t == next_collect) {
      if (a.out_z) {
#pragma unroll
        for (int k = 0; k < D; ++k) a.out_z[(sidx * D + k) * C + c] = s.x[k];
      }
      if (a.out_pe) a.out_pe[sidx * C + c] = s.U;
      ++sidx;
      next_collect += a.thinning;
    }
  }
}

// The whole per-chain launch body.
template <class Model, typename R, bool ADAPT, bool EXTERNAL>
AMCMC_HD void arwmh_chain_run(const Model& m, const StateView<R>& st, const RunView<R>& a, int64_t c) {
  constexpr int D = Model::D;
  ChainRegs<R, D> s;
  load_chain(s, st, c);
  const Philox rng(a.seed, (uint64_t)(c + a.chain_offset));
  arwmh_chain_range<Model, R, ADAPT, EXTERNAL>(s, m, a, rng, st.C, c, 0, a.n_steps);
  store_chain<R, D, ADAPT>(s, st, c);
}

#ifdef __CUDACC__
template <class Model, typename R, bool ADAPT, bool EXTERNAL>
__global__ void __launch_bounds__(64, (sizeof(R) == 4 ? 7 : 1))
arwmh_small_kernel(const Model m, const StateView<R> st, const RunView<R> a) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= st.C) return;
  arwmh_chain_run<Model, R, ADAPT, EXTERNAL>(m, st, a, c);
}

// ---------------------------------------------------------------------------------------------
// Few chains: two warps per 32 chains.  With one warp per scheduler a step is bound by that warp's own in-order issue
// (~700 dependent instructions, 0.93 us per step from 1 to ~5,000 chains), not by any pipe.  Here a second warp of the CTA,
// on another scheduler, generates the Philox / Box-Muller draws of step t + 1 into a shared-memory ring while the first
// runs the step t; one block barrier per step hands the buffers over.  Same draws, same arithmetic: every output is equal
// to arwmh_small_kernel's.
// ---------------------------------------------------------------------------------------------
template <class Model, typename R, bool ADAPT>
__global__ void __launch_bounds__(64) arwmh_small_duo_kernel(const Model m, const StateView<R> st, const RunView<R> a) {
  constexpr int D = Model::D;
  __shared__ R ring[2 * (D + 1) * 32];
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const int64_t cc = c < st.C ? c : st.C - 1;  // lanes past the end repeat the last chain (barriers stay uniform)
  const Philox rng(a.seed, (uint64_t)(cc + a.chain_offset));
  if (threadIdx.x >= 32) {  // producer
    for (int64_t t = 0; t < a.n_steps; ++t) {
      R z[D], u;
      philox_draws<R, D>(rng, (uint64_t)(a.i0 + t), z, u);
      R* buf = ring + (size_t)(t & 1) * (D + 1) * 32 + lane;
#pragma unroll
      for (int k = 0; k < D; ++k) buf[k * 32] = z[k];
      buf[D * 32] = u;
      __syncthreads();
    }
    return;
  }
  RunView<R> a2 = a;
  a2.ring = ring;
  ChainRegs<R, D> s;
  load_chain(s, st, cc);
  arwmh_chain_range<Model, R, ADAPT, false, true>(s, m, a2, rng, st.C, cc, 0, a.n_steps);
  if (c < st.C) store_chain<R, D, ADAPT>(s, st, c);
}

// ---------------------------------------------------------------------------------------------
// Balanced variant: one CTA of 12 worker warps per SM; the warps take (chain group, segment of steps) items from a queue.
//
// A warp of arwmh_small_kernel owns its 32 chains for the whole launch, so the number of warps per scheduler is fixed by the
// chain count: 65,536 chains on 148 SMs are 13.8 warps per SM = 4 + 4 + 3 + 3 per scheduler, and the launch lasts as long as
// the schedulers with 4 (measured, chain-steps/s at 1 / 2 / 3 / 3.5 / 4 warps per scheduler: 2.60 / 3.63 / 3.74 / 3.33 / 3.80
// x 10^10 -- a scheduler is saturated from 2 warps on, and 65,536 chains take exactly as long as 75,776).  Here every SM runs
// 12 worker warps (3 per scheduler, always busy) over its ~14 groups of 32 chains: a worker pops a group, runs `seg` steps
// with the state in registers, parks the registers (bit for bit, no LDL^T <-> Cholesky conversion) in the group's
// shared-memory slot and pushes the group back.  All four schedulers stay saturated until the SM's work is done.  The
// trajectories do not depend on the schedule (counter RNG, raw register hand-off).  Workers per SM, measured at 65,536 chains:
// 8 (255 registers) 3.55e10, 12 (168 registers, no spills) 3.82e10, 16 (128 registers) 3.77e10 chain-steps/s; the ASSS step
// (asss_small.cuh) needs more registers: 8 workers 7.85e9, 12 workers 7.64e9, plain kernel 6.88e9.
// ---------------------------------------------------------------------------------------------
#ifndef AMCMC_BAL_WARPS
#define AMCMC_BAL_WARPS 12
#endif
constexpr int kBalWarps = AMCMC_BAL_WARPS;
#ifndef AMCMC_BAL_WARPS_ASSS
#define AMCMC_BAL_WARPS_ASSS 8
#endif
constexpr int kBalWarpsAsss = AMCMC_BAL_WARPS_ASSS;
constexpr int kBalQueue = 32;  // ring capacity >= groups per CTA

template <typename R, int D> struct ChainSlot {
  static constexpr int NREG = 3 * D + ChainRegs<R, D>::NL + 4;
  static AMCMC_HD void save(const ChainRegs<R, D>& s, R* slot, int lane) {
    int r = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) slot[(r++) * 32 + lane] = s.x[k];
#pragma unroll
    for (int k = 0; k < D; ++k) slot[(r++) * 32 + lane] = s.mu[k];
#pragma unroll
    for (int k = 0; k < D; ++k) slot[(r++) * 32 + lane] = s.Dg[k];
#pragma unroll
    for (int k = 0; k < ChainRegs<R, D>::NL; ++k) slot[(r++) * 32 + lane] = s.Lt[k];
    slot[(r++) * 32 + lane] = s.U;
    slot[(r++) * 32 + lane] = s.lam;
    slot[(r++) * 32 + lane] = s.macc;
    slot[(r++) * 32 + lane] = s.asc;
  }
  static AMCMC_HD void restore(ChainRegs<R, D>& s, const R* slot, int lane) {
    int r = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) s.x[k] = slot[(r++) * 32 + lane];
#pragma unroll
    for (int k = 0; k < D; ++k) s.mu[k] = slot[(r++) * 32 + lane];
#pragma unroll
    for (int k = 0; k < D; ++k) s.Dg[k] = slot[(r++) * 32 + lane];
#pragma unroll
    for (int k = 0; k < ChainRegs<R, D>::NL; ++k) s.Lt[k] = slot[(r++) * 32 + lane];
    s.U = slot[(r++) * 32 + lane];
    s.lam = slot[(r++) * 32 + lane];
    s.macc = slot[(r++) * 32 + lane];
    s.asc = slot[(r++) * 32 + lane];
  }
};

// range runner of the ARWMH kernel (asss_small.cuh has the ASSS one)
template <class Model, typename R, bool ADAPT, bool EXTERNAL> struct ArwmhRange {
  static AMCMC_HD void run(ChainRegs<R, Model::D>& s, const Model& m, const RunView<R>& a, const Philox& rng, int64_t C, int64_t c,
                           int64_t t0, int64_t t1) {
    arwmh_chain_range<Model, R, ADAPT, EXTERNAL>(s, m, a, rng, C, c, t0, t1);
  }
};

template <class Model, typename R, bool ADAPT, class Runner, int WARPS = kBalWarps>
__global__ void __launch_bounds__(32 * WARPS, 1)
arwmh_small_balanced_kernel(const Model m, const StateView<R> st, const RunView<R> a, int64_t n_groups, int seg) {
  constexpr int D = Model::D;
  using Slot = ChainSlot<R, D>;
  extern __shared__ __align__(16) unsigned char bal_smem[];
  __shared__ int q_ring[kBalQueue];
  __shared__ int q_head, q_tail, q_left;
  __shared__ int64_t t_done[kBalQueue];
  R* slots = reinterpret_cast<R*>(bal_smem);
  // this CTA's groups: a contiguous range, the remainder spread over the first CTAs
  const int64_t base = n_groups / gridDim.x, rem = n_groups % gridDim.x;
  const int64_t g0 = (int64_t)blockIdx.x * base + ((int64_t)blockIdx.x < rem ? blockIdx.x : rem);
  const int ng = (int)(base + ((int64_t)blockIdx.x < rem ? 1 : 0));
  if (threadIdx.x == 0) {
    for (int k = 0; k < kBalQueue; ++k) { q_ring[k] = k < ng ? k : -1; t_done[k] = 0; }
    q_head = 0;
    q_tail = ng;
    q_left = ng;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t C = st.C, T = a.n_steps;
  for (;;) {
    int item = -1;
    if (lane == 0) {
      const long long w0 = clock64();
      for (;;) {
        if (*(volatile int*)&q_left == 0) break;
        const int h = *(volatile int*)&q_head;
        if (h < *(volatile int*)&q_tail) {
          if (atomicCAS(&q_head, h, h + 1) == h) {
            do { item = atomicExch(&q_ring[h % kBalQueue], -1); } while (item < 0);  // the pusher publishes right after reserving
            break;
          }
        } else {
          __nanosleep(200);  // (exponential back-off to 2 us was measured: 17.60 ms instead of 17.39 ms -- parked groups wait longer)
          if (clock64() - w0 > 200000000000LL) __trap();  // ~100 s: a scheduling bug must end as an error, not as a hung GPU
        }
      }
    }
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item < 0) break;
    const int64_t c = (g0 + item) * 32 + lane;
    const int64_t t0 = t_done[item];
    const int64_t t1 = t0 + seg < T ? t0 + seg : T;
    R* slot = slots + (size_t)item * (Slot::NREG * 32);
    if (c < C) {
      ChainRegs<R, D> s;
      if (t0 == 0) load_chain(s, st, c);
      else Slot::restore(s, slot, lane);
      const Philox rng(a.seed, (uint64_t)(c + a.chain_offset));
      Runner::run(s, m, a, rng, C, c, t0, t1);
      if (t1 == T) store_chain<R, D, ADAPT>(s, st, c);
      else Slot::save(s, slot, lane);
    }
    __syncwarp();
    if (lane == 0) {
      if (t1 == T) {
        atomicSub(&q_left, 1);
      } else {
        t_done[item] = t1;
        __threadfence_block();
        const int k = atomicAdd(&q_tail, 1);
        atomicExch(&q_ring[k % kBalQueue], item);
      }
    }
  }
}

// ARWMH.init (python/kernels/arwmh.py:111-136): q0 ~ U(-r, r)^d (unless given), U0, loc = q0, scale = I, ...
template <class Model, typename R>
__global__ void arwmh_small_init_kernel(const Model m, const StateView<R> st, uint64_t seed, int64_t chain_offset,
                                        R radius, int use_given_z) {
  constexpr int D = Model::D;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t C = st.C;
  if (c >= C) return;
  R q[D];
  if (use_given_z) {
#pragma unroll
    for (int k = 0; k < D; ++k) q[k] = st.z[k * C + c];
  } else {
    const Philox rng(seed, (uint64_t)(c + chain_offset));
    constexpr int NBLK = (D + 3) / 4;
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
      uint32_t o[4];
      rng.block(kInitStep, (uint32_t)b, o);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (4 * b + k < D) q[4 * b + k] = (R)((word_to_uniform(o[k]) * 2.0f - 1.0f) * (float)radius);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) st.z[k * C + c] = q[k];
  }
  st.pe[c] = m.potential(q);
#pragma unroll
  for (int k = 0; k < D; ++k) st.loc[k * C + c] = q[k];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) st.scale[tri_full(i, j) * C + c] = (i == j) ? (R)1 : (R)0;
  st.lam[c] = 0;
  st.macc[c] = 0;
  st.asc[c] = 0;
}

template <class Model, typename R>
__global__ void potential_small_kernel(const Model m, int64_t n, const R* __restrict__ q, R* __restrict__ out) {
  constexpr int D = Model::D;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  R x[D];
#pragma unroll
  for (int k = 0; k < D; ++k) x[k] = q[k * n + c];
  out[c] = m.potential(x);
}
#endif  // __CUDACC__

}  // namespace amcmc
