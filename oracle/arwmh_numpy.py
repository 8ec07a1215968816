"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE) -- NumPy restatement of the
reference's adaptive random-walk Metropolis hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``adaptive_mcmc_b200``) never imports it and has no CPU fallback.

PARITY STATUS: **parity unpinned** in the strict sense -- the reference
(savelovme/adaptive-mcmc) is pure Python on JAX + NumPyro, neither of which is
installable in this image, and the reference ships no tests / golden vectors for
this path.  What this oracle *is* pinned against (tests/test_oracle.py):
  * the energy-scale pins and posterior table recorded in the reference's
    notebooks,
  * independent SciPy log-pdf formulas for every potential,
  * the algebraic contract of ``cholesky_update`` (L'L'^T = LL^T + c xx^T),
  * Random123 known-answer vectors for Philox4x32-10,
  * (on the GPU side, tests/test_gpu_quality.py) the reference's recorded 100-seed quality table
    rmse_means / wasserstein / mmd for ARWMH and ASSS, reproduced at the reference's own run lengths.

Every function cites the reference file:line it restates (paths relative to
/root/reference).  Third-party arithmetic that is *not* under /root/reference
(NumPyro, unpinned in python/environment.yml:10) is restated from its published
algorithm and marked "3rd-party".

Conventions: chain-major arrays ``[C, d]`` (C independent chains); ``dt`` is
np.float32 (the reference's precision: x64 is never enabled) or np.float64.
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np

# --------------------------------------------------------------------------
# State records -- python/kernels/arwmh.py:15-28 (same field names and order)
# --------------------------------------------------------------------------
ARWMHState = namedtuple(
    "ARWMHState",
    ["i", "z", "potential_energy", "mean_accept_prob", "adapt_state", "as_change", "rng_key"],
)
ARWMHAdaptState = namedtuple("ARWMHAdaptState", ["loc", "scale", "log_step_size"])

LOG_2PI_HALF = 0.5 * math.log(2.0 * math.pi)

# eight_schools data literal: python/jupyter/posteriordb_eight-schools.ipynb:L502-503
EIGHT_SCHOOLS_Y = np.array([28, 8, -3, 7, -1, 1, 18, 12], dtype=np.float64)
EIGHT_SCHOOLS_SIGMA = np.array([15, 10, 16, 11, 9, 11, 10, 18], dtype=np.float64)


# --------------------------------------------------------------------------
# Potentials  U(q) = -log p(q, data) - log|det J|   (unconstrained space)
# --------------------------------------------------------------------------
def potential_eight_schools(q, y=EIGHT_SCHOOLS_Y, sigma=EIGHT_SCHOOLS_SIGMA):
    """python/scripts/run_eight_schools_lr_decay.py:26-35 (non-centred model).

    q = [mu, log tau, theta_base[0..7]] (sorted-site ravel order).
    mu ~ N(0,5); tau ~ HalfCauchy(5) with exp-transform Jacobian;
    theta_base ~ N(0,1); y_j ~ N(mu + tau*theta_base_j, sigma_j).
    """
    dt = q.dtype
    y = np.asarray(y, dt)
    sigma = np.asarray(sigma, dt)
    mu, t, eta = q[..., 0], q[..., 1], q[..., 2:]
    tau = np.exp(t)
    lp_mu = -0.5 * (mu / dt.type(5)) ** 2 - dt.type(math.log(5.0) + LOG_2PI_HALF)
    # HalfCauchy(5).log_prob(tau) + log|d tau / d t|   (3rd-party: NumPyro HalfCauchy)
    lp_tau = dt.type(math.log(2.0) - math.log(math.pi) - math.log(5.0)) - np.log1p((tau / dt.type(5)) ** 2) + t
    lp_eta = np.sum(-0.5 * eta * eta - dt.type(LOG_2PI_HALF), axis=-1)
    theta = mu[..., None] + tau[..., None] * eta
    r = (y - theta) / sigma
    lp_obs = np.sum(-0.5 * r * r - np.log(sigma) - dt.type(LOG_2PI_HALF), axis=-1)
    return (-(lp_mu + lp_tau + lp_eta + lp_obs)).astype(dt)


def _student_t_logpdf(x, df, loc, scale, dt):
    """3rd-party: numpyro.distributions.StudentT.log_prob."""
    yv = (x - dt.type(loc)) / dt.type(scale)
    zc = (
        math.log(scale)
        + 0.5 * math.log(df)
        + 0.5 * math.log(math.pi)
        + math.lgamma(0.5 * df)
        - math.lgamma(0.5 * (df + 1.0))
    )
    return -dt.type(0.5 * (df + 1.0)) * np.log1p(yv * yv / dt.type(df)) - dt.type(zc)


def diamonds_center(X):
    """python/scripts/run_diamonds_lr_decay.py:26-29: Xc = X[:,1:] - colmean."""
    X = np.asarray(X)
    return X[:, 1:] - X[:, 1:].mean(axis=0)


def potential_diamonds(q, X, Y):
    """python/scripts/run_diamonds_lr_decay.py:24-40.

    q = [Intercept, b[0..Kc-1], log sigma].  b ~ N(0,1); Intercept ~ StudentT(3,8,10);
    sigma ~ Folded StudentT(3,0,10) with exp Jacobian; Y ~ N(Intercept + Xc b, sigma).
    """
    dt = q.dtype
    Xc = diamonds_center(np.asarray(X, np.float64)).astype(dt)
    Y = np.asarray(Y, dt)
    N, Kc = Xc.shape
    icpt, b, s = q[..., 0], q[..., 1 : 1 + Kc], q[..., 1 + Kc]
    sig = np.exp(s)
    lp_b = np.sum(-0.5 * b * b - dt.type(LOG_2PI_HALF), axis=-1)
    lp_i = _student_t_logpdf(icpt, 3.0, 8.0, 10.0, dt)
    lp_s = dt.type(math.log(2.0)) + _student_t_logpdf(sig, 3.0, 0.0, 10.0, dt) + s
    mu = icpt[..., None] + b @ Xc.T
    r = (Y - mu) / sig[..., None]
    lp_y = np.sum(-0.5 * r * r, axis=-1) - dt.type(N) * (s + dt.type(LOG_2PI_HALF))
    return (-(lp_b + lp_i + lp_s + lp_y)).astype(dt)


def potential_kidiq(q, kid_score, mom_hs, mom_iq):
    """python/scripts/run_kidiq_kidscore_lr_decay.py:29-41.

    q = [beta0, beta1, beta2, log sigma]; flat prior on beta; sigma ~ HalfCauchy(2.5).
    """
    dt = q.dtype
    kid = np.asarray(kid_score, dt)
    hs = np.asarray(mom_hs, dt)
    iq = np.asarray(mom_iq, dt)
    b0, b1, b2, s = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    sig = np.exp(s)
    lp_s = dt.type(math.log(2.0) - math.log(math.pi) - math.log(2.5)) - np.log1p((sig / dt.type(2.5)) ** 2) + s
    mu = b0[..., None] + b1[..., None] * hs + b2[..., None] * iq
    r = (kid - mu) / sig[..., None]
    lp_y = np.sum(-0.5 * r * r, axis=-1) - dt.type(kid.shape[0]) * (s + dt.type(LOG_2PI_HALF))
    return (-(lp_s + lp_y)).astype(dt)


def potential_std_normal(q):
    """N(0, I_d) target used by python/jupyter/asumptions_check.ipynb cells 17-28
    (potential_fn = 0.5*x^2 up to the normalising constant)."""
    dt = q.dtype
    return (0.5 * np.sum(q * q, axis=-1)).astype(dt)


def potential_gaussian(q, prec_chol):
    """BASELINE.json config 5 (not in reference): N(0, Sigma), U = 0.5*||P^T q||^2
    with P = lower Cholesky factor of the precision matrix Sigma^-1 = P P^T."""
    dt = q.dtype
    v = q @ np.asarray(prec_chol, dt)  # rows: q^T P
    return (0.5 * np.sum(v * v, axis=-1)).astype(dt)


def make_potential(model, **data):
    if model == "eight_schools":
        y = data.get("y", EIGHT_SCHOOLS_Y)
        sigma = data.get("sigma", EIGHT_SCHOOLS_SIGMA)
        return lambda q: potential_eight_schools(q, y, sigma)
    if model == "diamonds":
        return lambda q: potential_diamonds(q, data["X"], data["Y"])
    if model == "kidiq":
        return lambda q: potential_kidiq(q, data["kid_score"], data["mom_hs"], data["mom_iq"])
    if model == "std_normal":
        return potential_std_normal
    if model == "gaussian":
        return lambda q: potential_gaussian(q, data["prec_chol"])
    raise ValueError(f"unknown model {model!r}")


# --------------------------------------------------------------------------
# cholesky_update -- 3rd-party numpyro.distributions.util.cholesky_update
# (call sites python/kernels/arwmh.py:190, asss.py:254)
# --------------------------------------------------------------------------
def cholesky_update(L, x, coef):
    """chol(L L^T + coef * x x^T), batched over the leading axis.

    Restated from NumPyro's published algorithm (Krause & Igel 2015, LDL^T
    rank-one recurrence): normalise to unit diagonal (L/diag, D=diag^2), scan over
    columns j:  gamma = b*D_j + coef*w_j^2;  D_j' = gamma/b;  b <- gamma/D_j;
    w <- w - w_j*L_j;  L_j <- L_j + (coef*w_j/gamma)*w;  result L~ * sqrt(D').
    L: [C,d,d], x: [C,d], coef: scalar or [C].  A zero diagonal yields NaN (0/0),
    which is what makes the reference keep the old factor at gamma == 1.
    """
    dt = L.dtype
    C, d, _ = L.shape
    coef = np.broadcast_to(np.asarray(coef, dt), (C,))
    with np.errstate(all="ignore"):
        diag = np.einsum("cii->ci", L).copy()
        Lu = L / diag[:, None, :]
        D = diag * diag
        b = np.ones(C, dt)
        w = x.astype(dt).copy()
        Dn = np.empty_like(D)
        Ln = np.empty_like(Lu)
        for j in range(d):
            wj = w[:, j].copy()
            Lj = Lu[:, :, j]
            gamma = b * D[:, j] + coef * wj * wj
            Dj_new = gamma / b
            b = gamma / D[:, j]
            w = w - wj[:, None] * Lj
            Ln[:, :, j] = Lj + (coef * wj / gamma)[:, None] * w
            Dn[:, j] = Dj_new
        return (Ln * np.sqrt(Dn)[:, None, :]).astype(dt)


# --------------------------------------------------------------------------
# Philox4x32-10 counter RNG (product RNG; the reference uses threefry2x32 which
# is only reproducible through the external-draws interface)
# --------------------------------------------------------------------------
_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011). All args uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, np.uint32).copy() for v in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.broadcast_to(np.asarray(k0, np.uint32), c0.shape).copy()
    k1 = np.broadcast_to(np.asarray(k1, np.uint32), c0.shape).copy()
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = k0 + _PH_W0
            k1 = k1 + _PH_W1
    return c0, c1, c2, c3


def philox_words(seed, chain_ids, step, n_words):
    """Word stream of the product RNG for (seed, chain, step).

    counter = (step_lo, (step_hi & 0xFFFFFF) | (blk << 24), chain_lo, chain_hi),
    key = (seed_lo, seed_hi); block ``blk`` yields words 4*blk .. 4*blk+3.
    Returns uint32 [C, n_words].  Mirrors adaptive_mcmc_b200/csrc/rng.cuh.
    """
    chain_ids = np.asarray(chain_ids, np.uint64)
    C = chain_ids.shape[0]
    nblk = (n_words + 3) // 4
    out = np.empty((C, nblk * 4), np.uint32)
    step = int(step)
    s_lo = np.uint32(step & 0xFFFFFFFF)
    s_hi = (step >> 32) & 0xFFFFFF
    for blk in range(nblk):
        c1 = np.uint32(s_hi | (blk << 24))
        w = philox4x32_10(
            np.full(C, s_lo, np.uint32),
            np.full(C, c1, np.uint32),
            (chain_ids & _MASK32).astype(np.uint32),
            (chain_ids >> np.uint64(32)).astype(np.uint32),
            np.uint32(seed & 0xFFFFFFFF),
            np.uint32((seed >> 32) & 0xFFFFFFFF),
        )
        for k in range(4):
            out[:, 4 * blk + k] = w[k]
    return out[:, :n_words]


def box_muller_words(words):
    """Box-Muller on consecutive word pairs (2k, 2k+1) in float32, exactly as the device does."""
    wa = words[:, 0::2]
    wb = words[:, 1::2]
    u1 = wa.astype(np.float32) * np.float32(2.0**-32) + np.float32(2.0**-33)
    th = np.float32(2.0 * math.pi) * (wb.astype(np.int32).astype(np.float32) * np.float32(2.0**-32))
    r = np.sqrt(np.float32(-2.0) * np.log(u1))
    z = np.empty((words.shape[0], 2 * wa.shape[1]), np.float32)
    z[:, 0::2] = r * np.cos(th)
    z[:, 1::2] = r * np.sin(th)
    return z


def words_to_draws(words, d, dt=np.float32):
    """Map Philox words to (normals[C,d], uniform[C]) exactly as the device does:
    Box-Muller on word pairs (2k, 2k+1) in float32:
       u1 = fma(float(w_a), 2^-32, 2^-33) in (0,1];  theta = 2*pi*(int32(w_b) * 2^-32)
       z_2k = r*cos(theta), z_2k+1 = r*sin(theta), r = sqrt(-2 ln u1)
    accept uniform = (w[2*ceil(d/2)] >> 8) * 2^-24 in [0,1).
    """
    npair = (d + 1) // 2
    wa = words[:, 0 : 2 * npair : 2]
    wb = words[:, 1 : 2 * npair : 2]
    u1 = wa.astype(np.float32) * np.float32(2.0**-32) + np.float32(2.0**-33)
    th = np.float32(2.0 * math.pi) * (wb.astype(np.int32).astype(np.float32) * np.float32(2.0**-32))
    r = np.sqrt(np.float32(-2.0) * np.log(u1))
    z = np.empty((words.shape[0], 2 * npair), np.float32)
    z[:, 0::2] = r * np.cos(th)
    z[:, 1::2] = r * np.sin(th)
    u = (words[:, 2 * npair] >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)
    return z[:, :d].astype(dt), u.astype(dt)


def n_words_for_dim(d):
    return 2 * ((d + 1) // 2) + 1


def philox_draws(seed, chain_ids, step, d, dt=np.float32):
    return words_to_draws(philox_words(seed, chain_ids, step, n_words_for_dim(d)), d, dt)


def philox_init_uniform(seed, chain_ids, d, radius=2.0, dt=np.float32):
    """q0 ~ U(-radius, radius)^d (NumPyro init_to_uniform radius 2; python/kernels/arwmh.py:44,111-115)
    drawn from the product RNG at the reserved step index 2^56-1."""
    w = philox_words(seed, chain_ids, (1 << 56) - 1, d)
    u = (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)
    return ((u * np.float32(2.0) - np.float32(1.0)) * np.float32(radius)).astype(dt)


# --------------------------------------------------------------------------
# ARWMH init / step / run -- python/kernels/arwmh.py:84-207
# --------------------------------------------------------------------------
def arwmh_init(potential, q0, rng_key=0):
    """python/kernels/arwmh.py:118-138: U0 = potential(q0); loc = q0; scale = I_d;
    log_step_size = 0; i = 0; mean_accept_prob = 0; as_change = 0."""
    q0 = np.asarray(q0)
    dt = q0.dtype
    C, d = q0.shape
    adapt = ARWMHAdaptState(
        loc=q0.copy(),
        scale=np.broadcast_to(np.eye(d, dtype=dt), (C, d, d)).copy(),
        log_step_size=np.zeros(C, dt),
    )
    return ARWMHState(
        i=0,
        z=q0.copy(),
        potential_energy=potential(q0).astype(dt),
        mean_accept_prob=np.zeros(C, dt),
        adapt_state=adapt,
        as_change=np.zeros(C, dt),
        rng_key=rng_key,
    )


def arwmh_step(
    state,
    potential,
    normals,
    uniforms,
    num_warmup=0,
    lr_decay=2.0 / 3.0,
    target_accept_prob=0.234,
    eps=1e-6,
    adapt=True,
):
    """One ARWMH.sample (python/kernels/arwmh.py:140-207) for C chains at once,
    with the step's draws supplied: normals [C,d] (prop_base, :165) and
    uniforms [C] (:174).  ``adapt=False`` is the frozen kernel of sample_Pnx
    (:230-249): position/energy move, everything else is reset/discarded.
    Returns (new_state, accept_prob[C], is_accepted[C])."""
    i = state.i
    z = state.z
    dt = z.dtype
    C, d = z.shape
    U = state.potential_energy
    mu_hat, L, lam = state.adapt_state
    one = dt.type(1)
    with np.errstate(all="ignore"):
        # :166-167  prop_scale = L*exp(lambda) + eps*I ; z' = z + prop_scale @ prop_base
        prop_scale = L * np.exp(lam)[:, None, None] + np.eye(d, dtype=dt) * dt.type(eps)
        z_prop = z + np.einsum("cij,cj->ci", prop_scale, normals.astype(dt))
        # :170-171
        U_prop = potential(z_prop).astype(dt)
        U_prop = np.where(np.isnan(U_prop), dt.type(np.inf), U_prop)
        # :173-174   jnp.clip(x, max=1) propagates NaN
        e = np.exp(U - U_prop)
        accept_prob = np.where(e > one, one, e).astype(dt)
        is_acc = uniforms.astype(dt) < accept_prob
        # :176-178
        z_new = np.where(is_acc[:, None], z_prop, z)
        U_new = np.where(is_acc, U_prop, U)
        # :180-185
        itr = i + 1
        n = itr if i < num_warmup else itr - num_warmup
        gamma = dt.type(1.0) / dt.type(n) ** dt.type(lr_decay)
        mean_acc_new = state.mean_accept_prob + (accept_prob - state.mean_accept_prob) / dt.type(n)
        if not adapt:
            new = ARWMHState(itr, z_new, U_new, mean_acc_new, state.adapt_state, state.as_change, state.rng_key)
            return new, accept_prob, is_acc
        # :188-191
        delta = z_new - mu_hat
        mu_new = mu_hat + gamma * delta
        chol = cholesky_update(np.sqrt(one - gamma) * L, delta, gamma)
        bad = np.isnan(chol).any(axis=(1, 2))
        L_new = np.where(bad[:, None, None], L, chol)
        # :193
        lam_new = lam + gamma * (accept_prob - dt.type(target_accept_prob))
        # :197  Frobenius norm
        diff = L_new * np.exp(lam_new)[:, None, None] - L * np.exp(lam)[:, None, None]
        as_change = np.sqrt(np.sum(diff * diff, axis=(1, 2))).astype(dt)
    new = ARWMHState(
        itr,
        z_new.astype(dt),
        U_new.astype(dt),
        mean_acc_new.astype(dt),
        ARWMHAdaptState(mu_new.astype(dt), L_new.astype(dt), lam_new.astype(dt)),
        as_change,
        state.rng_key,
    )
    return new, accept_prob, is_acc


def arwmh_run(
    state,
    potential,
    n_steps,
    draws=None,
    seed=0,
    chain_offset=0,
    thinning=1,
    collect_start=0,
    record_accept=False,
    **kw,
):
    """Drive ``n_steps`` steps.  ``draws`` = (normals[T,C,d], uniforms[T,C]) for the
    shared-draw parity mode, else the Philox stream keyed by (seed, global chain
    id, iteration index state.i).  Collection follows numpyro.util.fori_collect as
    used at python/utils/kernel_utils.py:29-32: sample k = state after
    collect_start + (k+1)*thinning steps.
    Returns (last_state, dict(z=[S,C,d], potential_energy=[S,C], accepts=[T,C]?))."""
    C, d = state.z.shape
    dt = state.z.dtype
    chain_ids = np.arange(C, dtype=np.uint64) + np.uint64(chain_offset)
    zs, pes, accs = [], [], []
    for t in range(n_steps):
        if draws is not None:
            nrm, uni = draws[0][t], draws[1][t]
        else:
            nrm, uni = philox_draws(seed, chain_ids, state.i, d, dt)
        state, _, acc = arwmh_step(state, potential, nrm, uni, **kw)
        if record_accept:
            accs.append(acc.copy())
        done = t + 1 - collect_start
        if done > 0 and done % thinning == 0:
            zs.append(state.z.copy())
            pes.append(state.potential_energy.copy())
    out = dict(
        z=np.stack(zs) if zs else np.zeros((0, C, d), dt),
        potential_energy=np.stack(pes) if pes else np.zeros((0, C), dt),
    )
    if record_accept:
        out["accepts"] = np.stack(accs)
    return state, out


# --------------------------------------------------------------------------
# RAM variant (NOT in the reference -- spec is ours)
# --------------------------------------------------------------------------
def ram_step(z, U, L, potential, normals, uniforms, n, lr_decay=2.0 / 3.0, target_accept_prob=0.234):
    """Robust adaptive Metropolis (Vihola 2012): x' = x + L z;  after accept/reject
    L'L'^T = L (I + eta_n (alpha - alpha*) z z^T / |z|^2) L^T  via a rank-1
    update (alpha > alpha*) or downdate (alpha < alpha*) with v = L z / |z|."""
    dt = z.dtype
    one = dt.type(1)
    with np.errstate(all="ignore"):
        Lz = np.einsum("cij,cj->ci", L, normals.astype(dt))
        z_prop = z + Lz
        U_prop = potential(z_prop).astype(dt)
        U_prop = np.where(np.isnan(U_prop), dt.type(np.inf), U_prop)
        e = np.exp(U - U_prop)
        alpha = np.where(e > one, one, e).astype(dt)
        acc = uniforms.astype(dt) < alpha
        z_new = np.where(acc[:, None], z_prop, z)
        U_new = np.where(acc, U_prop, U)
        # Vihola (2012): eta_n = min(1, d n^-lr_decay); eta_n <= 1 and alpha* < 1 keep
        # 1 + eta (alpha - alpha*) > 0, so the downdate stays positive definite
        eta = min(dt.type(1.0), dt.type(z.shape[1]) / dt.type(n) ** dt.type(lr_decay))
        coef = eta * (alpha - dt.type(target_accept_prob)) / np.sum(normals.astype(dt) ** 2, axis=1)
        chol = cholesky_update(L, Lz, coef)
        bad = np.isnan(chol).any(axis=(1, 2))
        L_new = np.where(bad[:, None, None], L, chol)
    return z_new.astype(dt), U_new.astype(dt), L_new.astype(dt), alpha, acc


def ram_run(z, U, L, potential, n_steps, draws, i0=0, record_accept=False, **kw):
    """n_steps of ram_step with supplied draws (normals[T,C,d], uniforms[T,C]); n = iteration + 1."""
    accs, zs = [], []
    for t in range(n_steps):
        z, U, L, _, acc = ram_step(z, U, L, potential, draws[0][t], draws[1][t], i0 + t + 1, **kw)
        zs.append(z.copy())
        if record_accept:
            accs.append(acc.copy())
    return z, U, L, dict(z=np.stack(zs), accepts=np.stack(accs) if record_accept else None)


# --------------------------------------------------------------------------
# Drivers -- python/utils/kernel_utils.py:8-12
# --------------------------------------------------------------------------
def ns_logscale(n_pow=6):
    """python/utils/kernel_utils.py:8-12."""
    return np.concatenate(
        [
            np.arange(0 if p < 1 else 10 ** (p - 1), 10**p, 10 ** (max(0, p - 2))) + 10 ** (max(0, p - 2))
            for p in range(n_pow + 1)
        ]
    )


# --------------------------------------------------------------------------
# Diagnostics -- 3rd-party numpyro.diagnostics (n_eff / r_hat of print_summary)
# --------------------------------------------------------------------------
def _fft_next_fast_len(target):
    # numpyro.diagnostics._fft_next_fast_len: smallest 2^a 3^b 5^c >= target
    if target <= 2:
        return target
    while True:
        m = target
        while m % 2 == 0:
            m //= 2
        while m % 3 == 0:
            m //= 3
        while m % 5 == 0:
            m //= 5
        if m == 1:
            return target
        target += 1


def autocorrelation(x, axis=0):
    x = np.asarray(x, np.float64)
    N = x.shape[axis]
    M = _fft_next_fast_len(N)
    M2 = 2 * M
    x = np.swapaxes(x, axis, -1)
    centered = x - x.mean(axis=-1, keepdims=True)
    freqvec = np.fft.rfft(centered, n=M2, axis=-1)
    gram = freqvec.real**2 + freqvec.imag**2
    autocorr = np.fft.irfft(gram, n=M2, axis=-1)[..., :N]
    autocorr = autocorr / np.arange(N, 0.0, -1)
    with np.errstate(invalid="ignore", divide="ignore"):
        autocorr = autocorr / autocorr[..., :1]
    return np.swapaxes(autocorr, axis, -1)


def autocovariance(x, axis=0):
    x = np.asarray(x, np.float64)
    return autocorrelation(x, axis) * x.var(axis=axis, keepdims=True)


def _compute_chain_variance_stable(x):
    chain_var = x.var(axis=1, ddof=1)
    var_within = chain_var.mean(axis=0)
    var_estimator = var_within * (x.shape[1] - 1) / x.shape[1]
    if x.shape[0] > 1:
        chain_mean = x.mean(axis=1)
        var_between = chain_mean.var(axis=0, ddof=1)
        var_estimator = var_estimator + var_between
    else:
        var_within = var_estimator
    return var_within, var_estimator


def gelman_rubin(x):
    x = np.asarray(x, np.float64)
    var_within, var_estimator = _compute_chain_variance_stable(x)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.sqrt(var_estimator / var_within)


def split_gelman_rubin(x):
    x = np.asarray(x, np.float64)
    N_half = x.shape[1] // 2
    return gelman_rubin(np.concatenate([x[:, :N_half], x[:, -N_half:]], axis=0))


def effective_sample_size(x):
    """x: [chains, draws, ...] -> n_eff[...] (Geyer initial monotone sequence)."""
    x = np.asarray(x, np.float64)
    assert x.ndim >= 2 and x.shape[1] >= 2
    gamma_k_c = autocovariance(x, axis=1)
    var_within, var_estimator = _compute_chain_variance_stable(x)
    rho_k = 1.0 - (var_within - gamma_k_c.mean(axis=0)) / var_estimator
    rho_k[0] = 1.0
    Rho_k = rho_k[:-1:2, ...] + rho_k[1::2, ...]
    Rho_init = Rho_k[:1]
    Rho_k = np.concatenate(
        [Rho_init, np.minimum.accumulate(np.clip(Rho_k[1:, ...], a_min=0, a_max=None), axis=0)], axis=0
    )
    tau = -1.0 + 2.0 * np.sum(Rho_k, axis=0)
    return np.prod(x.shape[:2]) / tau
