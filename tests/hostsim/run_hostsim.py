"""TEST-ONLY: drive tests/hostsim/libhostsim.so (host compilation of the CUDA kernel body)."""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def es_cst(sigma):
    h = 0.5 * math.log(2 * math.pi)
    return (math.log(5) + h - (math.log(2) - math.log(math.pi) - math.log(5)) + 8 * h
            + float(np.sum(np.log(sigma) + h)))


def run_es(state, n_steps, draws=None, seed=0, chain_offset=0, thinning=1, collect_start=0, num_warmup=0,
           lr_decay=2 / 3, target=0.234, eps=1e-6, adapt=True, y=None, sigma=None, kernel="arwmh", split=None):
    """split=seg: the run is cut into ranges of `seg` steps with the registers parked in a ChainSlot between them (the
    hand-off of the balanced launch)"""
    from oracle.arwmh_numpy import ARWMHAdaptState, ARWMHState, EIGHT_SCHOOLS_SIGMA, EIGHT_SCHOOLS_Y
    lib = C.CDLL(os.path.join(_HERE, "libhostsim.so"))
    y = EIGHT_SCHOOLS_Y if y is None else np.asarray(y, np.float64)
    sigma = EIGHT_SCHOOLS_SIGMA if sigma is None else np.asarray(sigma, np.float64)
    dt = state.z.dtype
    Cn, d = state.z.shape
    ii, jj = np.tril_indices(d)
    z = np.ascontiguousarray(state.z.T).copy()
    loc = np.ascontiguousarray(state.adapt_state.loc.T).copy()
    scale = np.ascontiguousarray(state.adapt_state.scale[:, ii, jj].T).copy()
    pe = state.potential_energy.copy(); asc = state.as_change.copy()
    macc = state.mean_accept_prob.copy() if hasattr(state, "mean_accept_prob") else np.zeros(Cn, dt)
    lam = state.adapt_state.log_step_size.copy() if hasattr(state.adapt_state, "log_step_size") else np.zeros(Cn, dt)
    S = max(0, (n_steps - collect_start) // thinning)
    out_z = np.zeros((S, d, Cn), dt); out_pe = np.zeros((S, Cn), dt); out_acc = np.zeros((n_steps, Cn), np.uint8)
    nrm = uni = None
    if draws is not None:
        nrm = np.ascontiguousarray(np.transpose(draws[0], (0, 2, 1)).astype(dt))
        uni = np.ascontiguousarray(draws[1].astype(dt)) if kernel != "asss" else \
            np.ascontiguousarray(np.transpose(draws[1], (0, 2, 1)).astype(dt))
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    extra = ()
    if split:
        f = lib.hostsim_es_split_f64 if dt == np.float64 else lib.hostsim_es_split_f32
        extra = (C.c_int(int(split)), C.c_int(1 if kernel == "asss" else 0))
    elif kernel == "asss":
        f = lib.hostsim_es_asss_f64 if dt == np.float64 else lib.hostsim_es_asss_f32
    else:
        f = lib.hostsim_es_f64 if dt == np.float64 else lib.hostsim_es_f32
    f.restype = None
    f(p(y), p(sigma), C.c_double(es_cst(sigma)), C.c_int64(Cn), p(z), p(pe), p(macc), p(loc), p(scale), p(lam), p(asc),
      C.c_int64(state.i), C.c_int64(n_steps), C.c_int64(thinning), C.c_int64(collect_start), C.c_int64(num_warmup),
      C.c_double(lr_decay), C.c_double(target), C.c_double(eps), C.c_uint64(seed), C.c_int64(chain_offset),
      p(nrm), p(uni), p(out_z), p(out_pe), p(out_acc), C.c_int(1 if adapt else 0), *extra)
    L = np.zeros((Cn, d, d), dt); L[:, ii, jj] = scale.T
    new = ARWMHState(state.i + n_steps, z.T.copy(), pe, macc, ARWMHAdaptState(loc.T.copy(), L, lam), asc, state.rng_key)
    return new, dict(z=np.transpose(out_z, (0, 2, 1)), potential_energy=out_pe, accepts=out_acc.astype(bool))
