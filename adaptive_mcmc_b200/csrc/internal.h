// internal.h -- model handle and per-family launcher prototypes (not part of the public ABI).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/amcmc.h"

struct amcmc_model {
  int model_id;
  int dtype;
  int dim;
  int device;
  int64_t n_rows;        // data rows (kidiq, diamonds)
  int n_arrays;
  void* d_arr[4];        // device copies of the model arrays, converted to `dtype`
  int64_t arr_len[4];
  double h_small[32];    // small host-side constants (eight_schools y/sigma, folded constants)
  double cst;            // folded additive constant of the potential
  // diamonds / tensor-core path extras (filled by the family that needs them)
  void* extra;           // family-private device/host block
  // scratch for amcmc_arwmh_run_host
  void* scratch;
  size_t scratch_bytes;
  cudaStream_t host_streams[2];  // [0] compute, [1] device->host sample copies (created lazily)
  cudaEvent_t host_events[4];    // [0..1] chunk computed, [2..3] chunk copied
  int host_streams_ready;
  // custom families (AMCMC_MODEL_CUSTOM): entry points of the plugin library the potential was compiled into
  void* plugin_handle;
  int (*plugin_run)(const amcmc_model*, const amcmc_state*, const amcmc_run_args*, cudaStream_t);
  int (*plugin_init)(const amcmc_model*, const amcmc_state*, uint64_t, int64_t, double, int, cudaStream_t);
  int (*plugin_potential)(const amcmc_model*, int64_t, const void*, void*, cudaStream_t);
};

namespace amcmc {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// Per-family entry points: run / init / potential.  Return amcmc_status.
#define AMCMC_DECL_FAMILY(name)                                                                              \
  int run_##name(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s);       \
  int init_##name(const amcmc_model* m, const amcmc_state* st, uint64_t seed, int64_t chain_offset,           \
                  double radius, int use_given_z, cudaStream_t s);                                            \
  int potential_##name(const amcmc_model* m, int64_t n, const void* q, void* out, cudaStream_t s);

AMCMC_DECL_FAMILY(std_normal)
AMCMC_DECL_FAMILY(eight_schools)
AMCMC_DECL_FAMILY(kidiq)

// diamonds: block-per-chain CUDA-core path (exact) -- block_diamonds.cu
int create_diamonds(amcmc_model* m, const double* X, int64_t n, int K, const double* Y);
int run_diamonds_block(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s);
int init_diamonds(const amcmc_model* m, const amcmc_state* st, uint64_t seed, int64_t chain_offset, double radius,
                  int use_given_z, cudaStream_t s);
int potential_diamonds_block(const amcmc_model* m, int64_t n, const void* q, void* out, cudaStream_t s);
// gaussian: block-per-chain ARWMH (d <= 32) and RAM (d <= 256) -- block_gaussian.cu
int create_gaussian(amcmc_model* m, const double* P, int d);
int run_gaussian(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s);
int init_gaussian(const amcmc_model* m, const amcmc_state* st, uint64_t seed, int64_t chain_offset, double radius,
                  int use_given_z, cudaStream_t s);
int potential_gaussian(const amcmc_model* m, int64_t n, const void* q, void* out, cudaStream_t s);
// diamonds: tcgen05 tensor-core path (shared adaptation state) -- diamonds_tc.cu
int create_diamonds_tc(amcmc_model* m, const double* X, int64_t n, int K, const double* Y);
void destroy_diamonds_tc(amcmc_model* m);
bool diamonds_tc_available(const amcmc_model* m);
int run_diamonds_tc_adapt(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a, cudaStream_t s);

}  // namespace amcmc
