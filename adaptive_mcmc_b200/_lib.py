"""ctypes binding of libamcmc.so (the C ABI in include/amcmc.h).

There is deliberately NO fallback: if the CUDA library is missing or fails to load
the import raises, and every entry point raises on a non-zero status.  PyTorch is
used only for device memory and streams (tensor.data_ptr(), current stream).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libamcmc.so")

AMCMC_F32, AMCMC_F64 = 0, 1
RNG_PHILOX, RNG_EXTERNAL = 0, 1
KERNEL_ARWMH, KERNEL_RAM, KERNEL_ASSS = 0, 1, 2
MODEL_STD_NORMAL, MODEL_EIGHT_SCHOOLS, MODEL_KIDIQ, MODEL_DIAMONDS, MODEL_GAUSSIAN, MODEL_CUSTOM = 0, 1, 2, 3, 4, 5
IMPL_AUTO, IMPL_REGISTER, IMPL_BLOCK, IMPL_TENSOR, IMPL_REGISTER_BALANCED = 0, 1, 2, 3, 4


class AmcmcState(C.Structure):
    """struct amcmc_state (include/amcmc.h)."""

    _fields_ = [
        ("n_chains", C.c_int64),
        ("dim", C.c_int32),
        ("dtype", C.c_int32),
        ("i", C.c_int64),
        ("z", C.c_void_p),
        ("potential_energy", C.c_void_p),
        ("mean_accept_prob", C.c_void_p),
        ("loc", C.c_void_p),
        ("scale", C.c_void_p),
        ("log_step_size", C.c_void_p),
        ("as_change", C.c_void_p),
    ]


class AmcmcRunArgs(C.Structure):
    """struct amcmc_run_args (include/amcmc.h)."""

    _fields_ = [
        ("n_steps", C.c_int64),
        ("thinning", C.c_int64),
        ("collect_start", C.c_int64),
        ("num_warmup", C.c_int64),
        ("lr_decay", C.c_double),
        ("target_accept_prob", C.c_double),
        ("eps", C.c_double),
        ("adapt", C.c_int32),
        ("rng_mode", C.c_int32),
        ("seed", C.c_uint64),
        ("chain_offset", C.c_int64),
        ("normals", C.c_void_p),
        ("uniforms", C.c_void_p),
        ("out_z", C.c_void_p),
        ("out_potential_energy", C.c_void_p),
        ("out_accept", C.c_void_p),
        ("kernel_kind", C.c_int32),
        ("impl", C.c_int32),
    ]


class AmcmcPooled(C.Structure):
    """struct amcmc_pooled (include/amcmc.h)."""

    _fields_ = [
        ("dim", C.c_int32),
        ("dtype", C.c_int32),
        ("loc", C.c_void_p),
        ("scale", C.c_void_p),
        ("log_step_size", C.c_void_p),
        ("cov", C.c_void_p),
        ("window", C.c_int64),
    ]


class AmcmcError(RuntimeError):
    pass


_lib = None

# every symbol include/amcmc.h declares (tests check that the .so exports all of them)
EXPORTED_SYMBOLS = (
    "amcmc_model_create",
    "amcmc_model_create_custom",
    "amcmc_model_destroy",
    "amcmc_model_dim",
    "amcmc_model_dtype",
    "amcmc_arwmh_init",
    "amcmc_arwmh_run",
    "amcmc_potential",
    "amcmc_arwmh_run_host",
    "amcmc_arwmh_init_host",
    "amcmc_pooled_run",
    "amcmc_pooled_stats",
    "amcmc_pooled_update",
    "amcmc_selftest_umma",
    "amcmc_host_chunk_samples",
    "amcmc_eval_kernel_sum",
    "amcmc_eval_mmd_sums",
    "amcmc_eval_sqdist_median",
    "amcmc_eval_cost_matrix",
    "amcmc_eval_moment",
    "amcmc_eval_assignment",
    "amcmc_eval_sinkhorn",
    "amcmc_jax_draws",
    "amcmc_last_error",
    "amcmc_version",
)


def lib():
    """Load libamcmc.so once.  Raises if it has not been built (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AmcmcError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C adaptive_mcmc_b200/csrc`.  There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    L.amcmc_last_error.restype = C.c_char_p
    L.amcmc_version.restype = C.c_int
    L.amcmc_model_create.restype = C.c_int
    L.amcmc_model_create.argtypes = [
        C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int,
        C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int64),
    ]
    L.amcmc_model_create_custom.restype = C.c_int
    L.amcmc_model_create_custom.argtypes = [
        C.POINTER(C.c_void_p), C.c_char_p, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int64),
    ]
    L.amcmc_model_destroy.restype = C.c_int
    L.amcmc_model_destroy.argtypes = [C.c_void_p]
    L.amcmc_model_dim.restype = C.c_int
    L.amcmc_model_dim.argtypes = [C.c_void_p]
    L.amcmc_model_dtype.restype = C.c_int
    L.amcmc_model_dtype.argtypes = [C.c_void_p]
    L.amcmc_arwmh_init.restype = C.c_int
    L.amcmc_arwmh_init.argtypes = [
        C.c_void_p, C.POINTER(AmcmcState), C.c_uint64, C.c_int64, C.c_double, C.c_int, C.c_void_p,
    ]
    L.amcmc_arwmh_run.restype = C.c_int
    L.amcmc_arwmh_run.argtypes = [C.c_void_p, C.POINTER(AmcmcState), C.POINTER(AmcmcRunArgs), C.c_void_p]
    L.amcmc_potential.restype = C.c_int
    L.amcmc_potential.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.amcmc_arwmh_run_host.restype = C.c_int
    L.amcmc_arwmh_run_host.argtypes = [C.c_void_p, C.POINTER(AmcmcState), C.POINTER(AmcmcRunArgs)]
    L.amcmc_arwmh_init_host.restype = C.c_int
    L.amcmc_arwmh_init_host.argtypes = [C.c_void_p, C.POINTER(AmcmcState), C.c_uint64, C.c_int64, C.c_double, C.c_int]
    L.amcmc_pooled_run.restype = C.c_int
    L.amcmc_pooled_run.argtypes = [C.c_void_p, C.POINTER(AmcmcState), C.POINTER(AmcmcPooled), C.POINTER(AmcmcRunArgs), C.c_void_p]
    L.amcmc_pooled_stats.restype = C.c_int
    L.amcmc_pooled_stats.argtypes = [C.POINTER(AmcmcState), C.POINTER(AmcmcPooled), C.c_void_p, C.c_void_p]
    L.amcmc_pooled_update.restype = C.c_int
    L.amcmc_pooled_update.argtypes = [C.POINTER(AmcmcPooled), C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    L.amcmc_selftest_umma.restype = C.c_int
    L.amcmc_selftest_umma.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.amcmc_eval_kernel_sum.restype = C.c_int
    L.amcmc_eval_kernel_sum.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int,
                                        C.POINTER(C.c_double), C.c_void_p]
    L.amcmc_host_chunk_samples.restype = C.c_int64
    L.amcmc_host_chunk_samples.argtypes = [C.c_int64, C.c_int64]
    L.amcmc_eval_mmd_sums.restype = C.c_int
    L.amcmc_eval_mmd_sums.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_void_p]
    L.amcmc_eval_sqdist_median.restype = C.c_int
    L.amcmc_eval_sqdist_median.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double), C.c_void_p]
    L.amcmc_eval_cost_matrix.restype = C.c_int
    L.amcmc_eval_cost_matrix.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    L.amcmc_eval_moment.restype = C.c_int
    L.amcmc_eval_moment.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_void_p]
    L.amcmc_eval_sinkhorn.restype = C.c_int
    L.amcmc_eval_sinkhorn.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
    L.amcmc_jax_draws.restype = C.c_int
    L.amcmc_jax_draws.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.amcmc_eval_assignment.restype = C.c_int
    L.amcmc_eval_assignment.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().amcmc_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise AmcmcError(f"{what}: status {rc}: {msg}")
