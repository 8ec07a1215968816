"""ctypes front-end for oracle/liboracle.so (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Loads the C restatement of the reference hot path (oracle/arwmh_oracle.c) and
exposes it with the same state records as oracle/arwmh_numpy.py.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  PARITY STATUS: parity unpinned (see arwmh_numpy.py header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .arwmh_numpy import ARWMHAdaptState, ARWMHState, diamonds_center, EIGHT_SCHOOLS_SIGMA, EIGHT_SCHOOLS_Y

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODEL_IDS = {"std_normal": 0, "eight_schools": 1, "kidiq": 2, "diamonds": 3, "gaussian": 4}


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "arwmh_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
    return _LIB


def model_arrays(model, **data):
    """-> (model_id, d, n_rows, [a0, a1, a2] float64 arrays)."""
    e = np.zeros(0)
    if model == "std_normal":
        return 0, int(data["d"]), 0, [e, e, e]
    if model == "eight_schools":
        y = np.asarray(data.get("y", EIGHT_SCHOOLS_Y), np.float64)
        s = np.asarray(data.get("sigma", EIGHT_SCHOOLS_SIGMA), np.float64)
        return 1, 10, 8, [y, s, e]
    if model == "kidiq":
        a = [np.ascontiguousarray(data[k], np.float64) for k in ("kid_score", "mom_hs", "mom_iq")]
        return 2, 4, a[0].shape[0], a
    if model == "diamonds":
        Xc = np.ascontiguousarray(diamonds_center(np.asarray(data["X"], np.float64)))
        Y = np.ascontiguousarray(data["Y"], np.float64)
        return 3, Xc.shape[1] + 2, Xc.shape[0], [Xc.ravel(), Y, e]
    if model == "gaussian":
        P = np.ascontiguousarray(data["prec_chol"], np.float64)
        return 4, P.shape[0], 0, [P.ravel(), e, e]
    raise ValueError(model)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _suffix(dt):
    return "_f64" if np.dtype(dt) == np.float64 else "_f32"


def potential(model, q, **data):
    q = np.ascontiguousarray(q)
    mid, d, n, arrs = model_arrays(model, **({"d": q.shape[1]} | data))
    out = np.empty(q.shape[0], q.dtype)
    f = getattr(lib(), "oracle_potential" + _suffix(q.dtype))
    f.restype = None
    f(C.c_int(mid), C.c_int(d), C.c_int64(n), _p(arrs[0]), _p(arrs[1]), _p(arrs[2]),
      C.c_int64(arrs[0].size), C.c_int64(arrs[1].size), C.c_int64(arrs[2].size),
      C.c_int64(q.shape[0]), _p(q), _p(out))
    return out


def init_uniform(seed, C_, d, radius=2.0, chain_offset=0, dt=np.float32):
    q0 = np.empty((C_, d), dt)
    f = getattr(lib(), "oracle_init_uniform" + _suffix(dt))
    f.restype = None
    f(C.c_uint64(seed), C.c_int64(chain_offset), C.c_int64(C_), C.c_int(d), C.c_double(radius), _p(q0))
    return q0


def arwmh_run(state, model, n_steps, draws=None, seed=0, chain_offset=0, thinning=1, collect_start=0,
              record_accept=False, num_warmup=0, lr_decay=2.0 / 3.0, target_accept_prob=0.234, eps=1e-6,
              adapt=True, n_threads=0, collect=True, **data):
    """Same contract as arwmh_numpy.arwmh_run, executed by the C oracle."""
    z = np.ascontiguousarray(state.z).copy()
    dt = z.dtype
    Cn, d = z.shape
    mid, d2, n, arrs = model_arrays(model, **({"d": d} | data))
    assert d2 == d
    pe = np.ascontiguousarray(state.potential_energy, dt).copy()
    macc = np.ascontiguousarray(state.mean_accept_prob, dt).copy()
    loc = np.ascontiguousarray(state.adapt_state.loc, dt).copy()
    scale = np.ascontiguousarray(state.adapt_state.scale, dt).copy()
    lam = np.ascontiguousarray(state.adapt_state.log_step_size, dt).copy()
    asc = np.ascontiguousarray(state.as_change, dt).copy()
    S = max(0, (n_steps - collect_start) // thinning) if collect else 0
    out_z = np.empty((S, Cn, d), dt) if S else None
    out_pe = np.empty((S, Cn), dt) if S else None
    out_acc = np.empty((n_steps, Cn), np.uint8) if record_accept else None
    nrm = uni = None
    if draws is not None:
        nrm = np.ascontiguousarray(draws[0], dt)
        uni = np.ascontiguousarray(draws[1], dt)
        assert nrm.shape == (n_steps, Cn, d) and uni.shape == (n_steps, Cn)
    f = getattr(lib(), "oracle_arwmh_run" + _suffix(dt))
    f.restype = C.c_int
    rc = f(C.c_int(mid), C.c_int(d), C.c_int64(n), _p(arrs[0]), _p(arrs[1]), _p(arrs[2]),
           C.c_int64(arrs[0].size), C.c_int64(arrs[1].size), C.c_int64(arrs[2].size),
           C.c_int64(Cn), _p(z), _p(pe), _p(macc), _p(loc), _p(scale), _p(lam), _p(asc),
           C.c_int64(state.i), C.c_int64(n_steps), _p(nrm), _p(uni), C.c_uint64(seed), C.c_int64(chain_offset),
           C.c_int64(thinning), C.c_int64(collect_start), _p(out_z), _p(out_pe), _p(out_acc),
           C.c_int64(num_warmup), C.c_double(lr_decay), C.c_double(target_accept_prob), C.c_double(eps),
           C.c_int(1 if adapt else 0), C.c_int(n_threads))
    assert rc == 0
    new = ARWMHState(state.i + n_steps, z, pe, macc, ARWMHAdaptState(loc, scale, lam), asc, state.rng_key)
    out = dict(
        z=out_z if S else np.zeros((0, Cn, d), dt),
        potential_energy=out_pe if S else np.zeros((0, Cn), dt),
    )
    if record_accept:
        out["accepts"] = out_acc.astype(bool)
    return new, out
