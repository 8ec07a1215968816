"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE) for the pooled-adaptation mode.

Pooled adaptation is NOT in the reference (BASELINE.json configs[3] asks for it); the spec is
ours (include/amcmc.h "Pooled adaptation") and this is its float64 restatement:
within a window every chain runs the frozen kernel of python/kernels/arwmh.py:161-178 (adapt_state
fixed, as sample_Pnx :230-249 does); between windows the reference's Robbins-Monro rule (:183-193)
is applied to the chain-average innovation.  PARITY STATUS: GPU-vs-our-oracle only.
"""
import numpy as np

from .arwmh_numpy import philox_draws


def pooled_init(z0):
    d = z0.shape[1]
    return dict(loc=z0.astype(np.float64).mean(0), L=np.eye(d), lam=0.0, cov=np.eye(d), window=0)


def pooled_window(potential, z, U, pool, K, i0, draws=None, seed=0, chain_offset=0, lr_decay=2 / 3,
                  target_accept_prob=0.234, eps=1e-6, adapt=True, record=False):
    """K frozen steps + one pooled update.  z [C,d] float64, U [C].  Returns (z, U, pool, info)."""
    C, d = z.shape
    z = z.copy(); U = U.copy()
    S = pool["L"] * np.exp(pool["lam"]) + eps * np.eye(d)
    acc_sum = np.zeros(C)
    accs, zs, pes = [], [], []
    chain_ids = np.arange(C, dtype=np.uint64) + np.uint64(chain_offset)
    with np.errstate(all="ignore"):
        for t in range(K):
            if draws is not None:
                nrm, uni = draws[0][t], draws[1][t]
            else:
                nrm, uni = philox_draws(seed, chain_ids, i0 + t, d, np.float64)
            zp = z + nrm @ S.T
            Up = potential(zp)
            Up = np.where(np.isnan(Up), np.inf, Up)
            e = np.exp(U - Up)
            alpha = np.where(e > 1.0, 1.0, e)
            acc = uni < alpha
            z = np.where(acc[:, None], zp, z)
            U = np.where(acc, Up, U)
            acc_sum += alpha
            if record:
                accs.append(acc.copy()); zs.append(z.copy()); pes.append(U.copy())
    info = dict(mean_accept=acc_sum / K)
    if record:
        info.update(accepts=np.stack(accs), z=np.stack(zs), potential_energy=np.stack(pes))
    if adapt:
        pool = dict(pool)
        n = pool["window"] + 1
        g = 1.0 / n ** lr_decay
        delta = z - pool["loc"]
        pool["loc"] = pool["loc"] + g * delta.mean(0)
        pool["cov"] = (1 - g) * pool["cov"] + g * (delta.T @ delta) / C
        pool["lam"] = pool["lam"] + g * (info["mean_accept"].mean() - target_accept_prob)
        try:
            pool["L"] = np.linalg.cholesky(pool["cov"])
        except np.linalg.LinAlgError:
            pass
        pool["window"] = n
    return z, U, pool, info
