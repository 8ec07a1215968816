"""Key handling for the reference's PRNG (jax.random, threefry2x32) on the host side of the drop-in: `PRNGKey(seed)` and
`split(key, n)` as plain NumPy, so that `ARWMH(..., rng="jax").init(rng_key, ...)` can derive the per-chain keys the way
NumPyro's `MCMC` does (one key per chain from `random.split(rng_key, num_chains)`; the key itself for a single chain,
python/kernels/arwmh.py:135) without JAX being installed.  The per-step stream itself is generated on the GPU
(csrc/jax_rng.cu, `amcmc_jax_draws`).  Algorithms: Threefry-2x32-20 (Salmon et al., SC'11) and jax._src.prng's
non-partitionable `threefry_2x32` / `_threefry_split_original`."""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _threefry2x32(key, x0, x1):
    k0, k1 = _U32(key[0]), _U32(key[1])
    with np.errstate(over="ignore"):
        ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
        x0 = np.asarray(x0, _U32).copy()
        x1 = np.asarray(x1, _U32).copy()
        x0 += ks[0]
        x1 += ks[1]
        for blk in range(5):
            for r in _ROT[blk & 1]:
                x0 += x1
                x1 = (x1 << _U32(r)) | (x1 >> _U32(32 - r))
                x1 ^= x0
            x0 += ks[(blk + 1) % 3]
            x1 += ks[(blk + 2) % 3] + _U32(blk + 1)
    return x0, x1


def prng_key(seed):
    """jax.random.PRNGKey(seed) -> uint32[2]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], _U32)


def split(key, num=2):
    """jax.random.split(key, num) -> uint32[num, 2]."""
    counts = np.arange(2 * num, dtype=_U32)
    y0, y1 = _threefry2x32(key, counts[:num], counts[num:])
    return np.concatenate([y0, y1]).reshape(num, 2)


def chain_keys(rng_key, num_chains):
    """Per-chain keys from whatever the caller passes as `rng_key`: an int seed, a uint32[2] key, or uint32[C, 2] keys."""
    k = np.asarray(rng_key)
    if k.ndim == 0:
        k = prng_key(int(k))
    k = k.astype(np.uint64).astype(_U32)
    if k.shape == (num_chains, 2):
        return k.copy()
    if k.shape != (2,):
        raise ValueError(f"rng_key must be an int seed, a uint32[2] key or uint32[{num_chains}, 2] keys; got shape {k.shape}")
    return k[None].copy() if num_chains == 1 else split(k, num_chains)
