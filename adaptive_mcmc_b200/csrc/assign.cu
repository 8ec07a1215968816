// assign.cu -- optimal one-to-one assignment on the GPU for `wasserstein_dist11_p`
// (reference: python/utils/evaluation.py:42-66, where scipy.optimize.linear_sum_assignment takes 20.7 s for the
// 10^4 x 10^4 cost matrix of one diamonds run, python/jupyter/posteriordb_diamonds.ipynb:L3342).
//
// Algorithm: Bertsekas' forward auction with epsilon-scaling, Jacobi form (every unassigned row bids in every round).
// Costs are quantised to 24-bit integers c_ij = rint(cost_ij * (2^24 - 1) / max cost) -- the resolution of the float32
// matrix itself -- and multiplied by (n + 1); with integer prices the final phase (epsilon = 1) then ends within
// n * epsilon < n + 1 = one cost quantum of the optimum, i.e. AT the optimum of the quantised problem, which is what the
// tests compare with SciPy's solver on the same integer matrix (equal optimal cost; the matchings may differ in ties).
//
// One round = three small kernels (bid / resolve / assign); HBM-bound: a bidding row streams its n int32 costs once
// (coalesced 16-byte loads) against the n prices, which stay in L2.  No host round trip inside a phase except the
// unassigned-rows counter, read back every few rounds.
#include <cooperative_groups.h>
#include <climits>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "internal.h"

namespace amcmc {

constexpr int kBidThreads = 256;
constexpr long long kNoBid = LLONG_MIN;

__global__ void assign_max_kernel(const float* __restrict__ c, int64_t total, unsigned int* __restrict__ out_bits) {
  unsigned int m = 0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const float v = c[k];
    const unsigned int b = (v > 0.f && v == v) ? __float_as_uint(v) : 0u;  // non-negative floats order like their bit patterns
    m = b > m ? b : m;
  }
  for (int o = 16; o; o >>= 1) { const unsigned int t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, m);
}

__global__ void assign_quantise_kernel(const float* __restrict__ c, int64_t total, const unsigned int* __restrict__ max_bits,
                                       int32_t* __restrict__ ci) {
  const float cmax = __uint_as_float(*max_bits);
  const double scale = cmax > 0.f ? 16777215.0 / (double)cmax : 0.0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const float v = c[k];
    ci[k] = (int32_t)llrint((double)(v > 0.f ? v : 0.f) * scale);
  }
}

struct Top2 {
  long long best, second;
  int arg;
};
__device__ __forceinline__ void top2_push(Top2& t, long long v, int j) {
  if (v > t.best || (v == t.best && j < t.arg)) { t.second = t.best; t.best = v; t.arg = j; }
  else if (v > t.second) t.second = v;
}
__device__ __forceinline__ void top2_merge(Top2& a, const Top2& b) {
  if (b.best > a.best || (b.best == a.best && b.arg < a.arg)) {
    a.second = a.best > b.second ? a.best : b.second;
    a.best = b.best; a.arg = b.arg;
  } else {
    a.second = a.second > b.best ? a.second : b.best;
  }
}

// one block per row; unassigned rows bid for their best column: price + (best - second best value) + eps
__global__ void __launch_bounds__(kBidThreads)
auction_bid_kernel(const int32_t* __restrict__ ci, int n, const long long* __restrict__ price, const int* __restrict__ col_of,
                   long long eps, long long* __restrict__ maxbid, long long* __restrict__ bid_val, int* __restrict__ bid_col) {
  const int i = blockIdx.x;
  if (col_of[i] >= 0) return;
  const long long np1 = (long long)n + 1;
  const int32_t* row = ci + (int64_t)i * n;
  Top2 t{kNoBid, kNoBid, INT_MAX};
  const int n4 = (n % 4 == 0 && (((uintptr_t)row) & 15) == 0) ? n / 4 : 0;
  for (int q = threadIdx.x; q < n4; q += kBidThreads) {
    const int4 c4 = reinterpret_cast<const int4*>(row)[q];
    const int j = 4 * q;
    top2_push(t, -(long long)c4.x * np1 - price[j], j);
    top2_push(t, -(long long)c4.y * np1 - price[j + 1], j + 1);
    top2_push(t, -(long long)c4.z * np1 - price[j + 2], j + 2);
    top2_push(t, -(long long)c4.w * np1 - price[j + 3], j + 3);
  }
  for (int j = 4 * n4 + threadIdx.x; j < n; j += kBidThreads) top2_push(t, -(long long)row[j] * np1 - price[j], j);
  for (int o = 16; o; o >>= 1) {
    Top2 u;
    u.best = __shfl_xor_sync(0xffffffffu, t.best, o);
    u.second = __shfl_xor_sync(0xffffffffu, t.second, o);
    u.arg = __shfl_xor_sync(0xffffffffu, t.arg, o);
    top2_merge(t, u);
  }
  __shared__ Top2 sh[kBidThreads / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBidThreads / 32; ++w) top2_merge(t, sh[w]);
    const long long gap = t.second == kNoBid ? 0 : t.best - t.second;  // n == 1: no competitor
    const long long bid = price[t.arg] + gap + eps;
    bid_val[i] = bid;
    bid_col[i] = t.arg;
    atomicMax(&maxbid[t.arg], bid);
  }
}

// the highest bid of a column wins; among equal bids the lowest row index
__global__ void auction_resolve_kernel(int n, const int* __restrict__ col_of, const long long* __restrict__ maxbid,
                                       const long long* __restrict__ bid_val, const int* __restrict__ bid_col, int* __restrict__ winner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || col_of[i] >= 0) return;
  const int j = bid_col[i];
  if (bid_val[i] == maxbid[j]) atomicMin(&winner[j], i);
}

__global__ void auction_assign_kernel(int n, int* __restrict__ col_of, int* __restrict__ row_of, long long* __restrict__ price,
                                      long long* __restrict__ maxbid, int* __restrict__ winner, int* __restrict__ n_unassigned) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int w = winner[j];
  if (w == INT_MAX) return;
  const int prev = row_of[j];
  if (prev >= 0) col_of[prev] = -1;       // the previous owner is outbid (it did not bid this round: it was assigned)
  else atomicSub(n_unassigned, 1);
  row_of[j] = w;
  col_of[w] = j;
  price[j] = maxbid[j];
  maxbid[j] = kNoBid;
  winner[j] = INT_MAX;
}

// Tail of a phase: once only a few rows are unassigned a Jacobi round is three kernel launches for one or two bids (93 % of
// all rounds at n = 2000 have at most 8 unassigned rows).  One CTA then finishes the phase in Gauss-Seidel order: pop an
// unassigned row, scan its n values with all threads, bid, assign at once, push the evicted owner -- ~2 us per bid
// instead of ~20 us per round.  The queue holds at most kTailQueue rows (the number of unassigned rows never grows).
constexpr int kTailThreads = 1024;
constexpr int kTailQueue = 256;
constexpr int kTailSmemCols = 18000;  // up to this many columns, prices and owners live in shared memory during the tail (216 KB)

// The tail works on reduced costs r_j = c_ij (n + 1) + price_j >= 0 (the value of column j is -r_j) as unsigned 64-bit
// integers: one IMAD.WIDE.U32 per element, and the two smallest are kept with selects (the lanes of a warp would diverge
// on almost every element with branches).  A thread visits its columns in increasing order, so a tie keeps the earlier one.
struct Low2 {
  unsigned long long best, second;
  int arg;
};
constexpr unsigned long long kNoCost = ~0ull;
__device__ __forceinline__ void low2_push(Low2& t, unsigned int c, unsigned int np1, unsigned long long price, int j) {
  const unsigned long long v = (unsigned long long)c * np1 + price;
  const bool lt = v < t.best;
  const unsigned long long s2 = v < t.second ? v : t.second;
  t.second = lt ? t.best : s2;
  t.best = lt ? v : t.best;
  t.arg = lt ? j : t.arg;
}
// 64-bit unsigned minimum over a warp with two 32-bit redux.sync instead of five shuffle levels of 64-bit values
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned h = __reduce_min_sync(0xffffffffu, hi);
  const unsigned l = __reduce_min_sync(0xffffffffu, hi == h ? lo : 0xffffffffu);
  return ((unsigned long long)h << 32) | l;
}
// the two smallest over the warp's 32 partial results; every lane returns the same (ties: lowest column index)
__device__ __forceinline__ Low2 warp_low2(const Low2& t) {
  Low2 r;
  r.best = warp_min_u64(t.best);
  r.arg = __reduce_min_sync(0xffffffffu, t.best == r.best ? t.arg : INT_MAX);
  const bool win = (t.best == r.best) && (t.arg == r.arg);  // exactly one lane: column indices are distinct (or all INT_MAX)
  r.second = warp_min_u64(win ? t.second : t.best);
  return r;
}

// SMEM_STATE: the CTA owns prices and column owners while it runs (nobody else bids), so they are staged in shared memory
// and written back at the end; a bid then only streams the row's n int32 costs, with all of a thread's 16-byte loads issued
// before the first comparison.  Measured per bid at n = 10^4 (clock64 build, -DAMCMC_ASSIGN_TIMING): 10.9k cycles with the
// warp results merged by shuffles and the 32 warps merged serially by thread 0 (6.0k scan + reduce, 4.7k merge + update);
// redux.sync merges at both levels bring it to the figure in profiles/r02_eval.md.
template <bool SMEM_STATE>
__global__ void __launch_bounds__(kTailThreads)
auction_tail_kernel(const int32_t* __restrict__ ci, int n, long long* price_g, volatile int* col_of, int* row_of_g,
                    long long eps, int* __restrict__ n_unassigned, long long max_bids, unsigned long long* __restrict__ bids_done) {
  extern __shared__ long long s_dyn[];
  long long* s_price = s_dyn;
  int* s_owner = reinterpret_cast<int*>(s_dyn + n);
  __shared__ int queue[kTailQueue];
  __shared__ int q_count;
  __shared__ Low2 sh[kTailThreads / 32];
  if (threadIdx.x == 0) q_count = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kTailThreads)
    if (col_of[i] < 0) {
      const int k = atomicAdd(&q_count, 1);
      if (k < kTailQueue) queue[k] = i;
    }
  __syncthreads();
  if (q_count > kTailQueue) return;  // too many for the tail: the caller goes on with Jacobi rounds
  if (SMEM_STATE)
    for (int j = threadIdx.x; j < n; j += kTailThreads) { s_price[j] = price_g[j]; s_owner[j] = row_of_g[j]; }
  long long* price = SMEM_STATE ? s_price : price_g;
  int* row_of = SMEM_STATE ? s_owner : row_of_g;
  __syncthreads();
  const unsigned int np1 = (unsigned int)n + 1u;
  const bool vec = (n % 4 == 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long done = 0;
#ifdef AMCMC_ASSIGN_TIMING
  long long tk[4] = {0, 0, 0, 0};
#define AT(k) { const long long c_ = clock64(); tk[k] += c_ - tl; tl = c_; }
#else
#define AT(k)
#endif
  int nq = q_count;
  while (nq > 0 && done < max_bids) {
#ifdef AMCMC_ASSIGN_TIMING
    long long tl = clock64();
#endif
    const int i = queue[nq - 1];
#ifdef AMCMC_ASSIGN_FIXROW  // experiment only (WRONG results): rows aliased onto 64 L2-resident ones, to separate HBM latency from issue time
    const int32_t* row = ci + (int64_t)(i & 63) * n;
#else
    const int32_t* row = ci + (int64_t)i * n;
#endif
    Low2 t{kNoCost, kNoCost, INT_MAX};
    if (vec) {
      const int4* row4 = reinterpret_cast<const int4*>(row);
      const int n4 = n >> 2;
      for (int q0 = threadIdx.x; q0 < n4; q0 += 4 * kTailThreads) {  // four independent 16-byte loads in flight per thread
        int4 c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = q0 + u * kTailThreads;
          c[u] = q < n4 ? __ldcs(row4 + q) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = q0 + u * kTailThreads;
          if (q < n4) {
            const int j = 4 * q;
            low2_push(t, (unsigned int)c[u].x, np1, (unsigned long long)price[j], j);
            low2_push(t, (unsigned int)c[u].y, np1, (unsigned long long)price[j + 1], j + 1);
            low2_push(t, (unsigned int)c[u].z, np1, (unsigned long long)price[j + 2], j + 2);
            low2_push(t, (unsigned int)c[u].w, np1, (unsigned long long)price[j + 3], j + 3);
          }
        }
      }
    } else {
      for (int j = threadIdx.x; j < n; j += kTailThreads) low2_push(t, (unsigned int)__ldg(row + j), np1, (unsigned long long)price[j], j);
    }
    t = warp_low2(t);
    AT(0)
    if (lane == 0) sh[warp] = t;
    __syncthreads();
    AT(1)
    if (warp == 0) {  // warp 0 merges the 32 warp results and applies the bid
      t = warp_low2(sh[lane]);
      if (lane == 0) {
        const int j = t.arg;
        const int prev = row_of[j];
        const long long gap = t.second == kNoCost ? 0 : (long long)(t.second - t.best);  // n == 1: no competitor
        price[j] = price[j] + gap + eps;
        row_of[j] = i;
        col_of[i] = j;
        if (prev >= 0) { col_of[prev] = -1; queue[nq - 1] = prev; }  // the evicted owner takes the slot of the popped row
        else { q_count = nq - 1; atomicSub(n_unassigned, 1); }
      }
    }
    AT(2)
    ++done;
    __syncthreads();
    nq = q_count;
    AT(3)
  }
#ifdef AMCMC_ASSIGN_TIMING
  if (threadIdx.x == 0 && done > 0)
    printf("[tail] bids %lld: own scan+reduce %lld, wait others %lld, merge+update %lld, sync %lld cycles per bid\n", done,
           tk[0] / done, tk[1] / done, tk[2] / done, tk[3] / done);
#endif
  if (SMEM_STATE)
    for (int j = threadIdx.x; j < n; j += kTailThreads) { price_g[j] = s_price[j]; row_of_g[j] = s_owner[j]; }
  if (threadIdx.x == 0) atomicAdd(bids_done, (unsigned long long)done);
}

// ---- the same tail on a thread-block cluster ------------------------------------------------------------------------------
// One SM issues ~0.45 instructions per cycle and scheduler on the dependent top-2 chains, so the single-CTA tail is bound by
// its own instruction stream (profiles/r02_eval.md).  Here a cluster of 8 CTAs splits the columns: CTA r keeps the prices of
// its slice in shared memory and scans only that slice of the bidder's row; the eight partial results are exchanged through
// distributed shared memory (every CTA writes its partial into every CTA's exchange slot, one cluster barrier per bid, two
// slot sets alternate) and every CTA merges them, so all CTAs carry identical copies of the queue and of the column owners
// and run the same control flow.  Rank 0 writes the global side (col_of, the unassigned counter).  10^4 x 10^4: 0.19 s against
// 0.23-0.25 s with the single CTA (~1.9 us per bid: HBM latency of the row, two block barriers, one cluster barrier);
// prefetching the rows of the eight local winners' owners into L2 during the exchange was measured: 0.23 s, not kept.
constexpr int kClu = 8;
constexpr int kCluThreads = 512;

__global__ void __cluster_dims__(kClu, 1, 1) __launch_bounds__(kCluThreads)
auction_tail_cluster_kernel(const int32_t* __restrict__ ci, int n, long long* price_g, int* col_of, int* row_of_g, long long eps,
                            int* __restrict__ n_unassigned, long long max_bids, unsigned long long* __restrict__ bids_done) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  extern __shared__ long long clu_dyn[];
  const int slice = (n + kClu - 1) / kClu;
  const int c0 = rank * slice, c1 = min(n, c0 + slice);
  long long* s_price = clu_dyn;                                   // [slice]
  int* s_owner = reinterpret_cast<int*>(clu_dyn + slice);         // [n] replica
  __shared__ Low2 xch[2][kClu];
  __shared__ Low2 sh[kCluThreads / 32];
  __shared__ int queue[kTailQueue];
  __shared__ int q_count;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the queue must be identical in every CTA: ordered compaction by one warp
  if (warp == 0) {
    int cnt = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const bool un = i < n && col_of[i] < 0;
      const unsigned m = __ballot_sync(0xffffffffu, un);
      if (un) {
        const int k = cnt + __popc(m & ((1u << lane) - 1u));
        if (k < kTailQueue) queue[k] = i;
      }
      cnt += __popc(m);
    }
    if (lane == 0) q_count = cnt;
  }
  for (int j = threadIdx.x; j < n; j += kCluThreads) s_owner[j] = row_of_g[j];
  for (int j = c0 + threadIdx.x; j < c1; j += kCluThreads) s_price[j - c0] = price_g[j];
  __syncthreads();
  int nq = q_count;
  if (nq > kTailQueue) return;  // every CTA takes the same decision (before any cluster barrier)
  cluster.sync();               // all exchange slots exist before anybody writes into them
  const unsigned int np1 = (unsigned int)n + 1u;
  long long done = 0;
  int par = 0;
  while (nq > 0 && done < max_bids) {
    const int i = queue[nq - 1];
    const int32_t* row = ci + (int64_t)i * n;
    Low2 t{kNoCost, kNoCost, INT_MAX};
    for (int j = c0 + threadIdx.x; j < c1; j += kCluThreads) low2_push(t, (unsigned int)__ldcs(row + j), np1, (unsigned long long)s_price[j - c0], j);
    t = warp_low2(t);
    if (lane == 0) sh[warp] = t;
    __syncthreads();
    if (warp == 0) {
      Low2 u = lane < kCluThreads / 32 ? sh[lane] : Low2{kNoCost, kNoCost, INT_MAX};
      u = warp_low2(u);
      if (lane < kClu) *cluster.map_shared_rank(&xch[par][rank], lane) = u;  // my partial into CTA `lane`
    }
    cluster.sync();
    if (warp == 0) {
      Low2 u = lane < kClu ? xch[par][lane] : Low2{kNoCost, kNoCost, INT_MAX};
      u = warp_low2(u);
      if (lane == 0) {
        const int j = u.arg;
        const int prev = s_owner[j];
        const long long gap = u.second == kNoCost ? 0 : (long long)(u.second - u.best);
        if (j >= c0 && j < c1) s_price[j - c0] += gap + eps;
        s_owner[j] = i;
        if (prev >= 0) queue[nq - 1] = prev;   // the evicted owner takes the slot of the popped row
        else q_count = nq - 1;
        if (rank == 0) {
          col_of[i] = j;
          if (prev >= 0) col_of[prev] = -1;
          else atomicSub(n_unassigned, 1);
        }
      }
    }
    ++done;
    par ^= 1;
    __syncthreads();
    nq = q_count;
  }
  cluster.sync();  // nobody leaves while a peer may still write into its exchange slots
  for (int j = c0 + threadIdx.x; j < c1; j += kCluThreads) price_g[j] = s_price[j - c0];
  if (rank == 0) {
    for (int j = threadIdx.x; j < n; j += kCluThreads) row_of_g[j] = s_owner[j];
    if (threadIdx.x == 0) atomicAdd(bids_done, (unsigned long long)done);
  }
}

__global__ void auction_reset_kernel(int n, int* __restrict__ col_of, int* __restrict__ row_of, long long* __restrict__ maxbid,
                                     int* __restrict__ winner, int* __restrict__ n_unassigned) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0) *n_unassigned = n;
  if (k >= n) return;
  col_of[k] = -1; row_of[k] = -1; maxbid[k] = kNoBid; winner[k] = INT_MAX;
}

__global__ void assign_total_kernel(const int32_t* __restrict__ ci, const float* __restrict__ c, int n, const int* __restrict__ col_of,
                                    unsigned long long* __restrict__ total_int, double* __restrict__ total_float) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long vi = 0;
  double vf = 0.0;
  if (i < n) {
    const int j = col_of[i];
    vi = (unsigned long long)ci[(int64_t)i * n + j];
    vf = (double)c[(int64_t)i * n + j];
  }
  for (int o = 16; o; o >>= 1) { vi += __shfl_xor_sync(0xffffffffu, vi, o); vf += __shfl_xor_sync(0xffffffffu, vf, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(total_int, vi); atomicAdd(total_float, vf); }
}

// per-device workspace that grows on demand (calls on one device are expected from one host thread at a time, as for eval.cu)
static char* assign_workspace(size_t need) {
  static char* buf[64] = {};
  static size_t cap[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (cap[dev] < need) {
    if (buf[dev]) cudaFree(buf[dev]);
    buf[dev] = nullptr;
    cap[dev] = 0;
    if (cudaMalloc(&buf[dev], need) != cudaSuccess) return nullptr;
    cap[dev] = need;
  }
  return buf[dev];
}

}  // namespace amcmc

using namespace amcmc;

// Optimal assignment of the n x n float32 cost matrix `cost` (device, row-major).  Outputs (all optional except col_of_row):
//   col_of_row   [n] int32 (device): column matched to every row
//   quantised    [n*n] int32 (device) or NULL: the integer matrix that was solved (for the parity tests)
//   out_host[3]  : sum of the float costs of the matching, sum of the quantised costs, number of auction rounds
extern "C" int amcmc_eval_assignment(const float* cost, int64_t n64, int32_t* col_of_row, int32_t* quantised, double* out_host,
                                     void* stream) {
  if (!cost || !col_of_row || n64 < 1 || n64 > 46000) { set_error("amcmc_eval_assignment: need 1 <= n <= 46000"); return AMCMC_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)n64;
  const int64_t total = (int64_t)n * n;
  int rc;
  char* buf = nullptr;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_ci = quantised ? 0 : al((size_t)total * 4);
  const size_t bytes = b_ci + 3 * al((size_t)n * 8) + 4 * al((size_t)n * 4) + 256;
  const bool cached = bytes <= ((size_t)1 << 30);  // up to 1 GiB stays allocated between calls (n <= ~16,000): a 400 MB
  if (cached) {                                    // cudaMalloc / cudaFree pair per call costs up to hundreds of ms at times
    buf = assign_workspace(bytes);
    if (!buf) { set_error("amcmc_eval_assignment: workspace allocation failed"); return AMCMC_ERR_CUDA; }
  } else if ((rc = check_cuda(cudaMalloc(&buf, bytes), "cudaMalloc(assignment workspace)"))) return rc;
  auto release = [&]() { if (!cached) cudaFree(buf); };
  char* p = buf;
  auto take = [&](size_t b) { char* q = p; p += al(b); return (void*)q; };
  int32_t* ci = quantised ? quantised : (int32_t*)take((size_t)total * 4);
  long long* price = (long long*)take((size_t)n * 8);
  long long* maxbid = (long long*)take((size_t)n * 8);
  long long* bid_val = (long long*)take((size_t)n * 8);
  int* row_of = (int*)take((size_t)n * 4);
  int* winner = (int*)take((size_t)n * 4);
  int* bid_col = (int*)take((size_t)n * 4);
  int* col_of = (int*)col_of_row;
  unsigned int* scal = (unsigned int*)take(64);  // [0] max bits, [2] unassigned counter, [4..5] int total, [6..7] float total
  int* n_un = (int*)(scal + 2);
  cudaMemsetAsync(scal, 0, 64, s);
  cudaMemsetAsync(price, 0, (size_t)n * 8, s);
  const int gs = 148 * 8;
  assign_max_kernel<<<gs, 256, 0, s>>>(cost, total, scal);
  assign_quantise_kernel<<<gs, 256, 0, s>>>(cost, total, scal, ci);
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (n <= kTailSmemCols && (size_t)n * 12 > 48 * 1024 &&
      (rc = check_cuda(cudaFuncSetAttribute(auction_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemCols * 12),
                       "cudaFuncSetAttribute(auction_tail_kernel)"))) { release(); return rc; }
  // cluster tail: from ~2,000 columns on (below, one CTA scans a row in a few hundred cycles anyway); AMCMC_ASSIGN_CLUSTER=0/1 overrides
  const size_t clu_smem = (size_t)((n + kClu - 1) / kClu) * 8 + (size_t)n * 4 + 16;
  static const int clu_env = [] { const char* e = getenv("AMCMC_ASSIGN_CLUSTER"); return e ? atoi(e) : -1; }();
  const bool use_cluster = (clu_env >= 0 ? clu_env != 0 : n >= 2048) && clu_smem <= 200 * 1024 && n >= kClu;
  if (use_cluster && clu_smem > 48 * 1024 &&
      (rc = check_cuda(cudaFuncSetAttribute(auction_tail_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)clu_smem),
                       "cudaFuncSetAttribute(auction_tail_cluster_kernel)"))) { release(); return rc; }
  // epsilon-scaling on costs multiplied by (n + 1): start at ~1/8 of the largest scaled cost, divide by 6 down to 1
  long long eps = ((long long)16777215 * (n + 1)) / 8;
  if (eps < 1) eps = 1;
  long rounds = 0, host_iters = 0;
  const bool dbg = getenv("AMCMC_ASSIGN_DEBUG") != nullptr;  // per-phase breakdown on stderr (scripts/probes/assign_perf.py)
  double ms_tail = 0, ms_jac = 0;
  long n_tail = 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (dbg) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
  unsigned long long* tail_bids = (unsigned long long*)(scal + 8);
  for (;;) {
    auction_reset_kernel<<<nb, 256, 0, s>>>(n, col_of, row_of, maxbid, winner, n_un);
    int un = n;
    while (un > 0) {
      const bool tail = un <= kTailQueue / 2;
      if (dbg) cudaEventRecord(e0, s);
      if (tail) {  // few rows left: finish the phase in one CTA (Gauss-Seidel order)
        if (use_cluster)
          auction_tail_cluster_kernel<<<kClu, kCluThreads, clu_smem, s>>>(ci, n, price, col_of, row_of, eps, n_un, (long long)1 << 22, tail_bids);
        else if (n <= kTailSmemCols)
          auction_tail_kernel<true><<<1, kTailThreads, (size_t)n * 12, s>>>(ci, n, price, col_of, row_of, eps, n_un, (long long)1 << 22, tail_bids);
        else
          auction_tail_kernel<false><<<1, kTailThreads, 0, s>>>(ci, n, price, col_of, row_of, eps, n_un, (long long)1 << 22, tail_bids);
      } else {
        for (int r = 0; r < 4; ++r) {
          auction_bid_kernel<<<n, kBidThreads, 0, s>>>(ci, n, price, col_of, eps, maxbid, bid_val, bid_col);
          auction_resolve_kernel<<<nb, 256, 0, s>>>(n, col_of, maxbid, bid_val, bid_col, winner);
          auction_assign_kernel<<<nb, 256, 0, s>>>(n, col_of, row_of, price, maxbid, winner, n_un);
        }
        rounds += 4;
      }
      if ((rc = check_cuda(cudaMemcpyAsync(&un, n_un, 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync"))) { release(); return rc; }
      if (dbg) cudaEventRecord(e1, s);
      if ((rc = check_cuda(cudaStreamSynchronize(s), "auction round"))) { release(); return rc; }
      if (dbg) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); if (tail) { ms_tail += ms; ++n_tail; } else ms_jac += ms; }
      if (++host_iters > 2000000) { release(); set_error("amcmc_eval_assignment: auction did not terminate"); return AMCMC_ERR_CUDA; }
    }
    if (dbg) {
      unsigned long long tb = 0;
      cudaMemcpy(&tb, tail_bids, 8, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[assign] eps %lld: jacobi rounds %ld (%.2f ms)  tail launches %ld bids %llu (%.2f ms)\n", eps, rounds, ms_jac, n_tail, tb, ms_tail);
    }
    if (eps == 1) break;
    eps /= 6;
    if (eps < 1) eps = 1;
  }
  if (out_host) {
    assign_total_kernel<<<nb, 256, 0, s>>>(ci, cost, n, col_of, (unsigned long long*)(scal + 4), (double*)(scal + 6));
    unsigned int h[8];
    if ((rc = check_cuda(cudaMemcpyAsync(h, scal, 32, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync"))) { release(); return rc; }
    if ((rc = check_cuda(cudaStreamSynchronize(s), "assignment totals"))) { release(); return rc; }
    unsigned long long ti; double tf;
    memcpy(&ti, h + 4, 8); memcpy(&tf, h + 6, 8);
    unsigned long long tb = 0;
    cudaMemcpy(&tb, tail_bids, 8, cudaMemcpyDeviceToHost);
    out_host[0] = tf; out_host[1] = (double)ti; out_host[2] = (double)rounds + (double)tb;  // Jacobi rounds + Gauss-Seidel bids
  }
  if (dbg) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
  rc = check_cuda(cudaGetLastError(), "amcmc_eval_assignment");
  release();
  return rc;
}
