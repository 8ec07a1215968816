// capi.cu -- the C ABI declared in include/amcmc.h: handle management, argument validation,
// family dispatch and the host-buffer entry point.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <dlfcn.h>
#include "internal.h"

namespace amcmc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return AMCMC_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return AMCMC_ERR_CUDA;
}

static size_t elt(int dtype) { return dtype == AMCMC_F64 ? 8 : 4; }

// host float64 -> device array of `dtype`
static int upload(const double* src, int64_t n, int dtype, void** out) {
  *out = nullptr;
  if (n <= 0) return AMCMC_OK;
  int rc = check_cuda(cudaMalloc(out, (size_t)n * elt(dtype)), "cudaMalloc(model array)");
  if (rc) return rc;
  if (dtype == AMCMC_F64) return check_cuda(cudaMemcpy(*out, src, (size_t)n * 8, cudaMemcpyHostToDevice), "cudaMemcpy");
  std::vector<float> tmp((size_t)n);
  for (int64_t i = 0; i < n; ++i) tmp[(size_t)i] = (float)src[i];
  return check_cuda(cudaMemcpy(*out, tmp.data(), (size_t)n * 4, cudaMemcpyHostToDevice), "cudaMemcpy");
}

static const double kLog2PiHalfH = 0.91893853320467274178;

}  // namespace amcmc

using namespace amcmc;

extern "C" {

const char* amcmc_last_error(void) { return g_err; }
int amcmc_version(void) { return AMCMC_VERSION; }

int amcmc_model_create(amcmc_model** out, int model_id, int dtype, int dim, int n_arrays,
                       const double* const* arrays, const int64_t* lens) {
  if (!out) { set_error("amcmc_model_create: out is NULL"); return AMCMC_ERR_ARG; }
  *out = nullptr;
  if (dtype != AMCMC_F32 && dtype != AMCMC_F64) { set_error("bad dtype %d", dtype); return AMCMC_ERR_ARG; }
  if (n_arrays < 0 || n_arrays > 4 || (n_arrays > 0 && (!arrays || !lens))) {
    set_error("bad n_arrays %d / NULL arrays", n_arrays);
    return AMCMC_ERR_ARG;
  }
  amcmc_model* m = (amcmc_model*)calloc(1, sizeof(amcmc_model));
  m->model_id = model_id;
  m->dtype = dtype;
  m->dim = dim;
  m->n_arrays = n_arrays;
  int rc = check_cuda(cudaGetDevice(&m->device), "cudaGetDevice");
  if (rc) { free(m); return rc; }
  switch (model_id) {
    case AMCMC_MODEL_STD_NORMAL:
      if (dim < 1) { set_error("std_normal: dim must be >= 1"); rc = AMCMC_ERR_ARG; }
      break;
    case AMCMC_MODEL_EIGHT_SCHOOLS: {
      if (n_arrays != 2 || lens[0] != 8 || lens[1] != 8 || dim != 10) {
        set_error("eight_schools: expects y[8], sigma[8], dim 10");
        rc = AMCMC_ERR_ARG;
        break;
      }
      double cst = std::log(5.0) + kLog2PiHalfH - (std::log(2.0) - std::log(M_PI) - std::log(5.0)) + 8 * kLog2PiHalfH;
      for (int j = 0; j < 8; ++j) {
        m->h_small[j] = arrays[0][j];
        m->h_small[8 + j] = arrays[1][j];
        if (!(arrays[1][j] > 0)) { set_error("eight_schools: sigma must be > 0"); rc = AMCMC_ERR_ARG; }
        cst += std::log(arrays[1][j]) + kLog2PiHalfH;
      }
      m->cst = cst;
      break;
    }
    case AMCMC_MODEL_KIDIQ: {
      if (n_arrays != 3 || lens[0] < 1 || lens[1] != lens[0] || lens[2] != lens[0] || dim != 4) {
        set_error("kidiq: expects kid_score[n], mom_hs[n], mom_iq[n], dim 4");
        rc = AMCMC_ERR_ARG;
        break;
      }
      m->n_rows = lens[0];
      m->cst = -(std::log(2.0) - std::log(M_PI) - std::log(2.5)) + (double)m->n_rows * kLog2PiHalfH;
      for (int k = 0; k < 3 && !rc; ++k) {
        rc = upload(arrays[k], lens[k], dtype, &m->d_arr[k]);
        m->arr_len[k] = lens[k];
      }
      break;
    }
    case AMCMC_MODEL_DIAMONDS: {
      if (n_arrays != 2 || lens[1] < 1 || lens[0] % lens[1] != 0 || lens[0] / lens[1] != dim - 1 || dim < 3 || dim > 32) {
        set_error("diamonds: expects X[n*K] (row-major, column 0 = ones), Y[n], dim = K + 1 <= 32");
        rc = AMCMC_ERR_ARG;
        break;
      }
      rc = create_diamonds(m, arrays[0], lens[1], dim - 1, arrays[1]);
      if (!rc) rc = create_diamonds_tc(m, arrays[0], lens[1], dim - 1, arrays[1]);
      break;
    }
    case AMCMC_MODEL_GAUSSIAN: {
      if (n_arrays != 1 || dim < 1 || lens[0] != (int64_t)dim * dim) {
        set_error("gaussian: expects P[d*d] (row-major lower Cholesky factor of the precision)");
        rc = AMCMC_ERR_ARG;
        break;
      }
      rc = create_gaussian(m, arrays[0], dim);
      break;
    }
    default:
      set_error("model id %d not available in this build", model_id);
      rc = AMCMC_ERR_UNSUPPORTED;
  }
  if (rc) { amcmc_model_destroy(m); return rc; }
  *out = m;
  return AMCMC_OK;
}

int amcmc_model_create_custom(amcmc_model** out, const char* plugin_path, int dtype, int n_arrays,
                              const double* const* arrays, const int64_t* lens) {
  if (!out) { set_error("amcmc_model_create_custom: out is NULL"); return AMCMC_ERR_ARG; }
  *out = nullptr;
  if (!plugin_path) { set_error("amcmc_model_create_custom: plugin_path is NULL"); return AMCMC_ERR_ARG; }
  if (dtype != AMCMC_F32 && dtype != AMCMC_F64) { set_error("bad dtype %d", dtype); return AMCMC_ERR_ARG; }
  if (n_arrays < 0 || n_arrays > 4 || (n_arrays > 0 && (!arrays || !lens))) {
    set_error("custom model: 0..4 data arrays");
    return AMCMC_ERR_ARG;
  }
  void* h = dlopen(plugin_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { set_error("custom model: dlopen(%s) failed: %s", plugin_path, dlerror()); return AMCMC_ERR_ARG; }
  typedef int (*int_fn)(void);
  int_fn f_dim = (int_fn)dlsym(h, "amcmc_plugin_dim");
  int_fn f_abi = (int_fn)dlsym(h, "amcmc_plugin_abi");
  void* f_run = dlsym(h, "amcmc_plugin_run");
  void* f_init = dlsym(h, "amcmc_plugin_init");
  void* f_pot = dlsym(h, "amcmc_plugin_potential");
  if (!f_dim || !f_abi || !f_run || !f_init || !f_pot) {
    set_error("custom model: %s does not export the amcmc_plugin_* entry points", plugin_path);
    dlclose(h);
    return AMCMC_ERR_ARG;
  }
  if (f_abi() != (int)sizeof(amcmc_model) * 1000 + AMCMC_VERSION) {
    set_error("custom model: %s was built against another version of the library (rebuild the plugin)", plugin_path);
    dlclose(h);
    return AMCMC_ERR_ARG;
  }
  amcmc_model* m = (amcmc_model*)calloc(1, sizeof(amcmc_model));
  m->model_id = AMCMC_MODEL_CUSTOM;
  m->dtype = dtype;
  m->dim = f_dim();
  m->n_arrays = n_arrays;
  m->plugin_handle = h;
  m->plugin_run = (decltype(m->plugin_run))f_run;
  m->plugin_init = (decltype(m->plugin_init))f_init;
  m->plugin_potential = (decltype(m->plugin_potential))f_pot;
  int rc = check_cuda(cudaGetDevice(&m->device), "cudaGetDevice");
  for (int k = 0; k < n_arrays && !rc; ++k) {
    if (lens[k] < 0 || (lens[k] > 0 && !arrays[k])) { set_error("custom model: bad array %d", k); rc = AMCMC_ERR_ARG; break; }
    if (lens[k] > 0) rc = upload(arrays[k], lens[k], dtype, &m->d_arr[k]);
    m->arr_len[k] = lens[k];
  }
  if (n_arrays > 0 && !rc) m->n_rows = lens[0];
  if (rc) { amcmc_model_destroy(m); return rc; }
  *out = m;
  return AMCMC_OK;
}

int amcmc_model_destroy(amcmc_model* m) {
  if (!m) return AMCMC_OK;
  if (m->plugin_handle) dlclose(m->plugin_handle);
  for (int k = 0; k < 4; ++k)
    if (m->d_arr[k]) cudaFree(m->d_arr[k]);
  if (m->scratch) cudaFree(m->scratch);
  if (m->model_id == AMCMC_MODEL_DIAMONDS) destroy_diamonds_tc(m);
  if (m->host_streams_ready) {
    for (int k = 0; k < 2; ++k) cudaStreamDestroy(m->host_streams[k]);
    for (int k = 0; k < 4; ++k) cudaEventDestroy(m->host_events[k]);
  }
  free(m);
  return AMCMC_OK;
}

int amcmc_model_dim(const amcmc_model* m) { return m ? m->dim : AMCMC_ERR_ARG; }
int amcmc_model_dtype(const amcmc_model* m) { return m ? m->dtype : AMCMC_ERR_ARG; }

static int validate_state(const amcmc_model* m, const amcmc_state* st, const char* who) {
  if (!m || !st) { set_error("%s: NULL model/state", who); return AMCMC_ERR_ARG; }
  if (st->dim != m->dim || st->dtype != m->dtype) {
    set_error("%s: state (dim %d, dtype %d) does not match model (dim %d, dtype %d)", who, st->dim, st->dtype, m->dim, m->dtype);
    return AMCMC_ERR_ARG;
  }
  if (st->n_chains < 1) { set_error("%s: n_chains must be >= 1", who); return AMCMC_ERR_ARG; }
  if (!st->z || !st->potential_energy || !st->mean_accept_prob || !st->loc || !st->scale || !st->log_step_size || !st->as_change) {
    set_error("%s: NULL state array", who);
    return AMCMC_ERR_ARG;
  }
  return AMCMC_OK;
}

int amcmc_arwmh_init(const amcmc_model* m, amcmc_state* st, uint64_t seed, int64_t chain_offset, double init_radius,
                     int use_given_z, void* stream) {
  int rc = validate_state(m, st, "amcmc_arwmh_init");
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  st->i = 0;
  switch (m->model_id) {
    case AMCMC_MODEL_STD_NORMAL: return init_std_normal(m, st, seed, chain_offset, init_radius, use_given_z, s);
    case AMCMC_MODEL_EIGHT_SCHOOLS: return init_eight_schools(m, st, seed, chain_offset, init_radius, use_given_z, s);
    case AMCMC_MODEL_KIDIQ: return init_kidiq(m, st, seed, chain_offset, init_radius, use_given_z, s);
    case AMCMC_MODEL_DIAMONDS: return init_diamonds(m, st, seed, chain_offset, init_radius, use_given_z, s);
    case AMCMC_MODEL_GAUSSIAN: return init_gaussian(m, st, seed, chain_offset, init_radius, use_given_z, s);
    case AMCMC_MODEL_CUSTOM: return m->plugin_init(m, st, seed, chain_offset, init_radius, use_given_z, s);
  }
  set_error("amcmc_arwmh_init: unsupported model %d", m->model_id);
  return AMCMC_ERR_UNSUPPORTED;
}

static int validate_run(const amcmc_model* m, const amcmc_state* st, const amcmc_run_args* a) {
  int rc = validate_state(m, st, "amcmc_arwmh_run");
  if (rc) return rc;
  if (!a) { set_error("amcmc_arwmh_run: NULL args"); return AMCMC_ERR_ARG; }
  if (a->n_steps < 0 || a->thinning < 1 || a->collect_start < 0 || a->num_warmup < 0) {
    set_error("amcmc_arwmh_run: need n_steps >= 0, thinning >= 1, collect_start >= 0, num_warmup >= 0");
    return AMCMC_ERR_ARG;
  }
  if (a->rng_mode == AMCMC_RNG_EXTERNAL && a->n_steps > 0 && (!a->normals || !a->uniforms)) {
    set_error("amcmc_arwmh_run: rng_mode EXTERNAL needs normals and uniforms");
    return AMCMC_ERR_ARG;
  }
  if (a->rng_mode != AMCMC_RNG_EXTERNAL && a->rng_mode != AMCMC_RNG_PHILOX) {
    set_error("amcmc_arwmh_run: bad rng_mode %d", a->rng_mode);
    return AMCMC_ERR_ARG;
  }
  if (!(a->lr_decay >= 0) || !(a->eps >= 0)) { set_error("amcmc_arwmh_run: lr_decay/eps must be >= 0"); return AMCMC_ERR_ARG; }
  return AMCMC_OK;
}

int amcmc_arwmh_run(const amcmc_model* m, amcmc_state* st, const amcmc_run_args* a, void* stream) {
  int rc = validate_run(m, st, a);
  if (rc) return rc;
  if (a->n_steps == 0) return AMCMC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool small_family = m->model_id == AMCMC_MODEL_STD_NORMAL || m->model_id == AMCMC_MODEL_EIGHT_SCHOOLS ||
                            m->model_id == AMCMC_MODEL_KIDIQ || m->model_id == AMCMC_MODEL_CUSTOM;
  if (a->kernel_kind != AMCMC_KERNEL_ARWMH && !(a->kernel_kind == AMCMC_KERNEL_RAM && m->model_id == AMCMC_MODEL_GAUSSIAN) &&
      !(a->kernel_kind == AMCMC_KERNEL_ASSS && (small_family || ((m->model_id == AMCMC_MODEL_DIAMONDS || m->model_id == AMCMC_MODEL_GAUSSIAN) && m->dim <= 32)))) {
    set_error("amcmc_arwmh_run: kernel kind %d not available for model %d", a->kernel_kind, m->model_id);
    return AMCMC_ERR_UNSUPPORTED;
  }
  switch (m->model_id) {
    case AMCMC_MODEL_STD_NORMAL: rc = run_std_normal(m, st, a, s); break;
    case AMCMC_MODEL_EIGHT_SCHOOLS: rc = run_eight_schools(m, st, a, s); break;
    case AMCMC_MODEL_KIDIQ: rc = run_kidiq(m, st, a, s); break;
    case AMCMC_MODEL_DIAMONDS: {
      // many chains, fp32, per-chain adaptation: the likelihood goes to the tensor cores (impl 3 forces, impl 2 forbids)
      // auto: the block kernel takes ~10 us per step with one chain per SM, 16 us with two and 31 us with three, the
      // tensor-core kernel a constant ~21 us up to 128 chains per SM: it wins from two chains per SM upwards
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const bool tc = a->adapt && a->kernel_kind == AMCMC_KERNEL_ARWMH && diamonds_tc_available(m) &&
                      (a->impl == 3 || (a->impl == 0 && st->n_chains > (int64_t)2 * sms));
      if (a->impl == 3 && !tc) { set_error("amcmc_arwmh_run: tensor-core path unavailable (needs fp32, K = 25, adapt = 1)"); rc = AMCMC_ERR_UNSUPPORTED; break; }
      rc = tc ? run_diamonds_tc_adapt(m, st, a, s) : run_diamonds_block(m, st, a, s);
      break;
    }
    case AMCMC_MODEL_GAUSSIAN: rc = run_gaussian(m, st, a, s); break;
    case AMCMC_MODEL_CUSTOM: rc = m->plugin_run(m, st, a, s); break;
    default:
      set_error("amcmc_arwmh_run: unsupported model %d", m->model_id);
      rc = AMCMC_ERR_UNSUPPORTED;
  }
  if (rc == AMCMC_OK) st->i += a->n_steps;
  return rc;
}

int amcmc_potential(const amcmc_model* m, int64_t n, const void* q, void* out, void* stream) {
  if (!m || !q || !out || n < 0) { set_error("amcmc_potential: bad argument"); return AMCMC_ERR_ARG; }
  if (n == 0) return AMCMC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  switch (m->model_id) {
    case AMCMC_MODEL_STD_NORMAL: return potential_std_normal(m, n, q, out, s);
    case AMCMC_MODEL_EIGHT_SCHOOLS: return potential_eight_schools(m, n, q, out, s);
    case AMCMC_MODEL_KIDIQ: return potential_kidiq(m, n, q, out, s);
    case AMCMC_MODEL_DIAMONDS: return potential_diamonds_block(m, n, q, out, s);
    case AMCMC_MODEL_GAUSSIAN: return potential_gaussian(m, n, q, out, s);
    case AMCMC_MODEL_CUSTOM: return m->plugin_potential(m, n, q, out, s);
  }
  set_error("amcmc_potential: unsupported model %d", m->model_id);
  return AMCMC_ERR_UNSUPPORTED;
}

// ARWMH.init with host buffers: device scratch, init kernel, copy everything back.
int amcmc_arwmh_init_host(amcmc_model* m, amcmc_state* hst, uint64_t seed, int64_t chain_offset, double init_radius,
                          int use_given_z) {
  int rc = validate_state(m, hst, "amcmc_arwmh_init_host");
  if (rc) return rc;
  const size_t w = elt(m->dtype);
  const int64_t C = hst->n_chains, d = hst->dim, np = d * (d + 1) / 2;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_vec = al((size_t)C * w), b_mat = al((size_t)C * d * w), b_tri = al((size_t)C * np * w);
  const size_t total = 4 * b_vec + 2 * b_mat + b_tri;
  if (m->scratch_bytes < total) {
    if (m->scratch) cudaFree(m->scratch);
    m->scratch = nullptr;
    m->scratch_bytes = 0;
    if ((rc = check_cuda(cudaMalloc(&m->scratch, total), "cudaMalloc(init_host scratch)"))) return rc;
    m->scratch_bytes = total;
  }
  char* p = (char*)m->scratch;
  auto take = [&](size_t b) { char* q = p; p += b; return (void*)q; };
  amcmc_state ds = *hst;
  ds.z = take(b_mat); ds.loc = take(b_mat); ds.scale = take(b_tri);
  ds.potential_energy = take(b_vec); ds.mean_accept_prob = take(b_vec);
  ds.log_step_size = take(b_vec); ds.as_change = take(b_vec);
  cudaStream_t s = 0;
  if (use_given_z && (rc = check_cuda(cudaMemcpyAsync(ds.z, hst->z, (size_t)C * d * w, cudaMemcpyHostToDevice, s), "H2D"))) return rc;
  if ((rc = amcmc_arwmh_init(m, &ds, seed, chain_offset, init_radius, use_given_z, (void*)s))) return rc;
  struct { void* h; void* d; size_t n; } cp[] = {
      {hst->z, ds.z, (size_t)C * d * w}, {hst->loc, ds.loc, (size_t)C * d * w}, {hst->scale, ds.scale, (size_t)C * np * w},
      {hst->potential_energy, ds.potential_energy, (size_t)C * w}, {hst->mean_accept_prob, ds.mean_accept_prob, (size_t)C * w},
      {hst->log_step_size, ds.log_step_size, (size_t)C * w}, {hst->as_change, ds.as_change, (size_t)C * w}};
  for (auto& c : cp)
    if ((rc = check_cuda(cudaMemcpyAsync(c.h, c.d, c.n, cudaMemcpyDeviceToHost, s), "D2H"))) return rc;
  if ((rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize"))) return rc;
  hst->i = 0;
  return AMCMC_OK;
}

// Host-buffer variant: H2D state (+ draws), fused run, D2H state + samples.  Pointers in
// *hst / *ha are host memory (pinned memory makes the copies truly asynchronous).  The run is cut
// into chunks of whole thinning periods; chunk k's samples travel device->host on a second stream
// while chunk k+1 computes, so the PCIe/C2C copy of the sample stream hides behind the kernel.
int64_t amcmc_host_chunk_samples(int64_t remaining, int64_t thinning) {
  if (remaining <= 0 || thinning < 1) return 0;
  const int64_t lo = (128 + thinning - 1) / thinning, hi = (2048 + thinning - 1) / thinning;
  if (remaining <= lo) return remaining;
  int64_t ns = (remaining + 1) / 2;
  if (ns < lo) ns = lo;
  if (ns > hi) ns = hi;
  if (remaining - ns < lo) ns = remaining;  // no stub chunk at the end
  return ns;
}

int amcmc_arwmh_run_host(amcmc_model* m, amcmc_state* hst, const amcmc_run_args* ha) {
  int rc = validate_run(m, hst, ha);
  if (rc) return rc;
  const size_t w = elt(m->dtype);
  const int64_t C = hst->n_chains, d = hst->dim, T = ha->n_steps;
  const int64_t np = d * (d + 1) / 2;
  const int64_t thin = ha->thinning;
  const int64_t S = (T > ha->collect_start) ? (T - ha->collect_start) / thin : 0;
  const bool ext = ha->rng_mode == AMCMC_RNG_EXTERNAL;
  // external draws per step: ARWMH / RAM normals[d], uniforms[1]; ASSS normals[d + 1], uniforms[52] (include/amcmc.h)
  const int64_t nrm_per_step = (ha->kernel_kind == AMCMC_KERNEL_ASSS) ? d + 1 : d;
  const int64_t uni_per_step = (ha->kernel_kind == AMCMC_KERNEL_ASSS) ? 52 : 1;
  const bool want_z = ha->out_z && S > 0, want_pe = ha->out_potential_energy && S > 0;
  // chunking in whole thinning periods: every chunk takes half of the samples that remain, between ~128 and ~2048 iterations
  // (amcmc_host_chunk_samples).  Long chunks first keep the launch boundaries few (each costs a state round trip through
  // HBM and a ramp-down of the kernel, ~45 us at 65,536 eight_schools chains); the copy of the LAST chunk cannot hide behind
  // a kernel, so the chunks taper: 41, 41, 41, 39, 19, 10, 5, 4 samples for 200 samples at thinning 50 (0.2 ms of exposed
  // copy and 8 launches, against 0.5 ms and 20 launches with uniform 512-iteration chunks).
  int64_t chunk_S = (want_z || want_pe) ? (2048 + thin - 1) / thin : 0;
  if (chunk_S > S) chunk_S = S;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_vec = al((size_t)C * w), b_mat = al((size_t)C * d * w), b_tri = al((size_t)C * np * w);
  const size_t b_oz = want_z ? al((size_t)chunk_S * d * C * w) : 0;
  const size_t b_ope = want_pe ? al((size_t)chunk_S * C * w) : 0;
  const size_t b_acc = ha->out_accept ? al((size_t)T * C) : 0;
  const size_t b_nrm = ext ? al((size_t)T * nrm_per_step * C * w) : 0;
  const size_t b_uni = ext ? al((size_t)T * uni_per_step * C * w) : 0;
  const size_t total = 4 * b_vec + 2 * b_mat + b_tri + 2 * (b_oz + b_ope) + b_acc + b_nrm + b_uni;
  if (m->scratch_bytes < total) {
    if (m->scratch) cudaFree(m->scratch);
    m->scratch = nullptr;
    m->scratch_bytes = 0;
    rc = check_cuda(cudaMalloc(&m->scratch, total), "cudaMalloc(run_host scratch)");
    if (rc) return rc;
    m->scratch_bytes = total;
  }
  if (!m->host_streams_ready) {
    for (int k = 0; k < 2; ++k)
      if ((rc = check_cuda(cudaStreamCreateWithFlags(&m->host_streams[k], cudaStreamNonBlocking), "cudaStreamCreate"))) return rc;
    for (int k = 0; k < 4; ++k)
      if ((rc = check_cuda(cudaEventCreateWithFlags(&m->host_events[k], cudaEventDisableTiming), "cudaEventCreate"))) return rc;
    m->host_streams_ready = 1;
  }
  cudaStream_t s0 = m->host_streams[0], s1 = m->host_streams[1];
  char* p = (char*)m->scratch;
  auto take = [&](size_t b) { char* q = p; p += b; return (void*)q; };
  amcmc_state ds = *hst;
  ds.z = take(b_mat); ds.loc = take(b_mat); ds.scale = take(b_tri);
  ds.potential_energy = take(b_vec); ds.mean_accept_prob = take(b_vec);
  ds.log_step_size = take(b_vec); ds.as_change = take(b_vec);
  void* d_oz[2] = {b_oz ? take(b_oz) : nullptr, b_oz ? take(b_oz) : nullptr};
  void* d_ope[2] = {b_ope ? take(b_ope) : nullptr, b_ope ? take(b_ope) : nullptr};
  uint8_t* d_acc = b_acc ? (uint8_t*)take(b_acc) : nullptr;
  void* d_nrm = b_nrm ? take(b_nrm) : nullptr;
  void* d_uni = b_uni ? take(b_uni) : nullptr;
#define H2D(dst, src, bytes) if ((rc = check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s0), "H2D"))) return rc
#define D2H(dst, src, bytes, st) if ((rc = check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st), "D2H"))) return rc
  H2D(ds.z, hst->z, (size_t)C * d * w);
  H2D(ds.loc, hst->loc, (size_t)C * d * w);
  H2D(ds.scale, hst->scale, (size_t)C * np * w);
  H2D(ds.potential_energy, hst->potential_energy, (size_t)C * w);
  H2D(ds.mean_accept_prob, hst->mean_accept_prob, (size_t)C * w);
  H2D(ds.log_step_size, hst->log_step_size, (size_t)C * w);
  H2D(ds.as_change, hst->as_change, (size_t)C * w);
  if (ext) {
    H2D(d_nrm, ha->normals, (size_t)T * nrm_per_step * C * w);
    H2D(d_uni, ha->uniforms, (size_t)T * uni_per_step * C * w);
  }
  int64_t t_done = 0, s_done = 0;
  int k = 0;
  while (t_done < T) {
    amcmc_run_args da = *ha;
    int64_t steps, ns;
    if (chunk_S > 0 && s_done < S) {
      ns = amcmc_host_chunk_samples(S - s_done, thin);
      da.collect_start = (t_done == 0) ? ha->collect_start : 0;
      steps = da.collect_start + ns * thin;
      if (s_done + ns == S) steps = T - t_done;  // the last chunk also takes the uncollected tail
    } else {
      ns = 0;
      steps = T - t_done;
      da.collect_start = steps;  // nothing left to collect
    }
    da.n_steps = steps;
    const int buf = k & 1;
    if (k >= 2) {  // buffer reuse: wait until its previous copy has drained
      if ((rc = check_cuda(cudaStreamWaitEvent(s0, m->host_events[2 + buf], 0), "cudaStreamWaitEvent"))) return rc;
    }
    da.out_z = (want_z && ns) ? d_oz[buf] : nullptr;
    da.out_potential_energy = (want_pe && ns) ? d_ope[buf] : nullptr;
    da.out_accept = d_acc ? d_acc + (size_t)t_done * C : nullptr;
    da.normals = ext ? (const void*)((const char*)d_nrm + (size_t)t_done * nrm_per_step * C * w) : nullptr;
    da.uniforms = ext ? (const void*)((const char*)d_uni + (size_t)t_done * uni_per_step * C * w) : nullptr;
    rc = amcmc_arwmh_run(m, &ds, &da, (void*)s0);
    if (rc) return rc;
    if (ns && (want_z || want_pe)) {
      if ((rc = check_cuda(cudaEventRecord(m->host_events[buf], s0), "cudaEventRecord"))) return rc;
      if ((rc = check_cuda(cudaStreamWaitEvent(s1, m->host_events[buf], 0), "cudaStreamWaitEvent"))) return rc;
      if (want_z) D2H((char*)ha->out_z + (size_t)s_done * d * C * w, d_oz[buf], (size_t)ns * d * C * w, s1);
      if (want_pe) D2H((char*)ha->out_potential_energy + (size_t)s_done * C * w, d_ope[buf], (size_t)ns * C * w, s1);
      if ((rc = check_cuda(cudaEventRecord(m->host_events[2 + buf], s1), "cudaEventRecord"))) return rc;
      ++k;
    }
    t_done += steps;
    s_done += ns;
  }
  D2H(hst->z, ds.z, (size_t)C * d * w, s0);
  D2H(hst->loc, ds.loc, (size_t)C * d * w, s0);
  D2H(hst->scale, ds.scale, (size_t)C * np * w, s0);
  D2H(hst->potential_energy, ds.potential_energy, (size_t)C * w, s0);
  D2H(hst->mean_accept_prob, ds.mean_accept_prob, (size_t)C * w, s0);
  D2H(hst->log_step_size, ds.log_step_size, (size_t)C * w, s0);
  D2H(hst->as_change, ds.as_change, (size_t)C * w, s0);
  if (b_acc) D2H(ha->out_accept, d_acc, (size_t)T * C, s0);
#undef H2D
#undef D2H
  if ((rc = check_cuda(cudaStreamSynchronize(s0), "cudaStreamSynchronize"))) return rc;
  if ((rc = check_cuda(cudaStreamSynchronize(s1), "cudaStreamSynchronize"))) return rc;
  hst->i = ds.i;
  return AMCMC_OK;
}

}  // extern "C"
