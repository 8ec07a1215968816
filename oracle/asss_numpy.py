"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE) -- NumPy restatement of the reference's
adaptive stereographic slice sampler, python/kernels/asss.py (a "next" row).

PARITY STATUS: parity unpinned (JAX/NumPyro cannot run here; see arwmh_numpy.py).  Pinned against the
eight_schools ASSS posterior agreement recorded in posteriordb_eight-schools.ipynb (cell 29) through
tests, and against itself (slice-sampler invariance on N(0, I)).

External-draws interface (shared-draw parity mode):
   normals[T, C, d+1]      v  (asss.py:231)
   uniforms[T, C, 2 + 50]  [0] = u_t (:236), [1] = theta_0 / 2pi (:61), [2+k] = k-th shrinkage draw (:84)
"""
import math
from collections import namedtuple

import numpy as np

from .arwmh_numpy import cholesky_update, philox_words, box_muller_words

ASSSState = namedtuple("ASSSState", ["i", "z", "potential_energy", "adapt_state", "as_change", "rng_key"])
ASSSAdaptState = namedtuple("ASSSAdaptState", ["loc", "scale"])
MAX_ITER = 50


def asss_init(potential, q0, rng_key=0):
    """asss.py:173-190: U0, loc = q0, scale = I, as_change = 0."""
    q0 = np.asarray(q0)
    C, d = q0.shape
    dt = q0.dtype
    return ASSSState(0, q0.copy(), potential(q0).astype(dt),
                     ASSSAdaptState(q0.copy(), np.broadcast_to(np.eye(d, dtype=dt), (C, d, d)).copy()),
                     np.zeros(C, dt), rng_key)


def _project(x, loc, S):
    """asss.py:33-44 (S lower-triangular, batched)."""
    y = np.stack([np.linalg.solve(S[c], x[c] - loc[c]) for c in range(x.shape[0])]).astype(x.dtype)
    nsq = np.sum(y * y, axis=1)
    return np.concatenate([2 * y / (nsq + 1)[:, None], ((nsq - 1) / (nsq + 1))[:, None]], axis=1)


def _inverse(z, loc, S):
    """asss.py:47-56."""
    xb = z[:, :-1] / (1 - z[:, -1])[:, None]
    return np.einsum("cij,cj->ci", S, xb) + loc


def asss_step(state, potential, normals, uniforms, num_warmup=0, lr_decay=2.0 / 3.0, eps=1e-6, adapt=True):
    """One ASSS.sample (asss.py:192-269) for C chains with supplied draws.
    Returns (new_state, n_shrink_iterations[C])."""
    i = state.i
    x = state.z
    dt = x.dtype
    C, d = x.shape
    loc, scale = state.adapt_state
    with np.errstate(all="ignore"):
        S = (scale + dt.type(eps) * np.eye(d, dtype=dt)) * dt.type(d) ** dt.type(0.5)  # :218

        def tpe(z):  # :222-225
            xx = _inverse(z, loc, S)
            return potential(xx).astype(dt) + dt.type(d) * np.log(1 - z[:, -1])

        z = _project(x, loc, S)  # :227
        pe_t = tpe(z)  # :228
        v = normals.astype(dt).copy()  # :231-233
        v -= np.sum(v * z, axis=1, keepdims=True) * z
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        t_pe = pe_t - np.log(uniforms[:, 0].astype(dt))  # :236-237
        # _shrinkage, asss.py:59-96
        theta = dt.type(2 * math.pi) * uniforms[:, 1].astype(dt)
        th_min = theta - dt.type(2 * math.pi)
        th_max = theta.copy()
        it = np.zeros(C, np.int64)
        active = np.ones(C, bool)
        for k in range(MAX_ITER + 1):
            z_th = z * np.cos(theta)[:, None] + v * np.sin(theta)[:, None]
            pe_th = tpe(z_th)
            pe_th = np.where(np.isnan(pe_th), dt.type(np.inf), pe_th)
            cont = (it < MAX_ITER) & ((pe_th > t_pe) | ((1 - z_th[:, -1]) < dt.type(eps)))
            active &= cont
            if not active.any():
                break
            th_min = np.where(active & (theta < 0), theta, th_min)
            th_max = np.where(active & (theta >= 0), theta, th_max)
            u = uniforms[:, 2 + min(k, MAX_ITER - 1)].astype(dt)
            theta = np.where(active, th_min + u * (th_max - th_min), theta)
            it = it + active
        theta = np.where(it >= MAX_ITER, dt.type(0), theta)  # :94
        z_new = z * np.cos(theta)[:, None] + v * np.sin(theta)[:, None]
        x_new = _inverse(z_new, loc, S).astype(dt)  # :241
        pe_new = potential(x_new).astype(dt)
        pe_new = np.where(np.isnan(pe_new), dt.type(np.inf), pe_new)
        # adaptation, :246-260
        itr = i + 1
        n = itr if i < num_warmup else itr - num_warmup
        gamma = dt.type(1.0) / dt.type(n) ** dt.type(lr_decay)
        delta = x_new - loc
        loc_new = loc + gamma * delta
        chol = cholesky_update(np.sqrt(dt.type(1) - gamma) * scale, delta, gamma)
        bad = np.isnan(chol).any(axis=(1, 2))
        scale_new = np.where(bad[:, None, None], scale, chol)
        as_change = np.linalg.norm(loc_new - loc, axis=1) + np.sqrt(np.sum((scale_new - scale) ** 2, axis=(1, 2)))
    if not adapt:  # sample_Pnx (asss.py:279-315): every step is taken with the GIVEN adapt_state, its update is dropped
        return ASSSState(itr, x_new, pe_new, state.adapt_state, state.as_change, state.rng_key), it
    new = ASSSState(itr, x_new, pe_new, ASSSAdaptState(loc_new.astype(dt), scale_new.astype(dt)), as_change.astype(dt),
                    state.rng_key)
    return new, it


def asss_words_to_draws(seed, chain_ids, step, d, dt=np.float32):
    """Philox stream of the product kernel for one ASSS step: d+1 normals from word pairs of blocks 0..,
    u_t and theta_0 from the next two words, shrinkage draw k from block 64 + k//4, word k%4."""
    npair = (d + 2) // 2
    base = 2 * npair
    w = philox_words(seed, chain_ids, step, base + 2)
    z = box_muller_words(w[:, :base])[:, : d + 1].astype(dt)
    C = len(chain_ids)
    uni = np.empty((C, 2 + MAX_ITER), dt)
    uni[:, 0] = ((w[:, base] >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)).astype(dt)
    uni[:, 1] = ((w[:, base + 1] >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)).astype(dt)
    ws = philox_words(seed, chain_ids, step, 4 * 64 + MAX_ITER + 2)[:, 4 * 64 : 4 * 64 + MAX_ITER]
    uni[:, 2:] = ((ws >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)).astype(dt)
    # u_t must be > 0 for log: the kernel uses (u + 2^-25)
    return z, uni


def asss_run(state, potential, n_steps, draws=None, seed=0, chain_offset=0, thinning=1, collect_start=0, **kw):
    C, d = state.z.shape
    dt = state.z.dtype
    ids = np.arange(C, dtype=np.uint64) + np.uint64(chain_offset)
    zs, pes, its = [], [], []
    for t in range(n_steps):
        if draws is not None:
            nrm, uni = draws[0][t], draws[1][t]
        else:
            nrm, uni = asss_words_to_draws(seed, ids, state.i, d, dt)
        state, it = asss_step(state, potential, nrm, uni, **kw)
        its.append(it)
        done = t + 1 - collect_start
        if done > 0 and done % thinning == 0:
            zs.append(state.z.copy()); pes.append(state.potential_energy.copy())
    return state, dict(z=np.stack(zs) if zs else np.zeros((0, C, d), dt),
                       potential_energy=np.stack(pes) if pes else np.zeros((0, C), dt), iterations=np.stack(its))
