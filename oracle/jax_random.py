"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE): the random stream of the reference.

The reference draws its proposal noise and accept uniforms from `jax.random` with the default
threefry2x32 PRNG (python/kernels/arwmh.py:162-165,174):

    rng_key, key_proposal, key_accept = random.split(rng_key, 3)          # :162
    prop_base = dist.Normal(0, 1).sample(key_proposal, (d,))               # :165  == random.normal(key, (d,))
    u         = dist.Uniform().sample(key_accept)                          # :174  == random.uniform(key, ())

JAX is a third-party dependency that is NOT under /root/reference and not installable in this image
(python/environment.yml:7 lists bare `jax`; the API usage implies 0.4.31 <= jax < 0.6).  This module restates
the published algorithms so that `rng_key -> (normals, uniforms)` can be reproduced without JAX:

  * Threefry-2x32, 20 rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123),
    pinned by the Random123 known-answer vectors that JAX's own test-suite uses (tests/random_test.py
    `testThreefry2x32`): see tests/test_jax_random.py.
  * `jax._src.prng`: `threefry_2x32(key, counts)` pairing (the counter array is split in two halves, an odd length is
    padded with one zero), `_threefry_split_original` (counts = iota(2 n)), `_threefry_random_bits_original`
    (counts = iota(n)) -- the NON-partitionable variant, the default before jax 0.5 (`jax_threefry_partitionable=False`).
  * `jax._src.random`: `uniform` (23 random mantissa bits | exponent of 1.0, minus 1, scaled, max(minval, .)) and
    `normal` (sqrt(2) * erfinv(uniform(-1 + ulp, 1))) in float32, with XLA's single-precision erfinv polynomial
    (M. Giles, "Approximating the erfinv function", 2012 -- the coefficients XLA's `ErfInv32` uses).

PARITY STATUS: the threefry block function is pinned by published vectors.  The JAX recipes are pinned by the values
the JAX documentation prints for `PRNGKey(0)` (split -> [4146024105 967050713] [2718843009 1272950319];
normal -> -0.20584226; uniform -> 0.41845703), which only come out right if every step above is right.  They were
recalled, not regenerated (no JAX here); scripts/jax_bridge.py re-checks them against the real library when it exists.
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(key, x0, x1):
    """Threefry-2x32-20 block function.  key: (k0, k1) uint32; x0, x1: uint32 arrays (counter words).
    Returns (y0, y1).  Random123 `threefry2x32_R(20, ctr, key)`."""
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
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(blk + 1) % 3]
            x1 += ks[(blk + 2) % 3] + _U32(blk + 1)
    return x0, x1


def prng_key(seed):
    """jax.random.PRNGKey(seed) for the threefry implementation: [seed >> 32, seed & 0xFFFFFFFF]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], _U32)


def threefry_2x32(key, counts):
    """jax._src.prng.threefry_2x32: hash a flat uint32 counter array; the array is cut in two halves that form the
    two words of each block (odd length: one zero appended, last output dropped)."""
    counts = np.asarray(counts, _U32).ravel()
    n = counts.size
    odd = n & 1
    if odd:
        counts = np.concatenate([counts, np.zeros(1, _U32)])
    h = counts.size // 2
    y0, y1 = threefry2x32(key, counts[:h], counts[h:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(key, num=2):
    """jax.random.split (non-partitionable threefry): keys [num, 2]."""
    return threefry_2x32(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def random_bits(key, n):
    """32-bit random words for a flat shape of n elements (non-partitionable threefry)."""
    return threefry_2x32(key, np.arange(n, dtype=_U32))


def uniform(key, shape=(), minval=0.0, maxval=1.0):
    """jax.random.uniform, float32."""
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    bits = random_bits(key, n)
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    out = np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)
    return out.reshape(shape)


def erfinv_f32(x):
    """XLA's float32 erf_inv (Giles 2012, single precision), evaluated in float32 like the device code."""
    x = np.asarray(x, np.float32)
    f = np.float32
    with np.errstate(all="ignore"):
        w = -np.log1p(-x * x).astype(np.float32)
        lt = w < f(5.0)
        wa = np.where(lt, w - f(2.5), np.sqrt(w) - f(3.0)).astype(np.float32)
        ca = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
              -0.00417768164, 0.246640727, 1.50140941)
        cb = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
              0.00943887047, 1.00167406, 2.83297682)
        p = np.where(lt, f(ca[0]), f(cb[0])).astype(np.float32)
        for a, b in zip(ca[1:], cb[1:]):
            p = (np.where(lt, f(a), f(b)) + p * wa).astype(np.float32)
        out = (p * x).astype(np.float32)
        out = np.where(np.abs(x) == f(1.0), np.copysign(f(np.inf), x), out)
    return out


def normal(key, shape=()):
    """jax.random.normal, float32: sqrt(2) * erfinv(u), u ~ U(nextafter(-1, 0), 1)."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(key, shape, lo, 1.0)
    return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32)


def arwmh_draws(rng_key, d, n_steps):
    """The draws ARWMH.sample consumes over n_steps steps from `rng_key` (python/kernels/arwmh.py:162-165,174).
    Returns (normals [T, d] float32, uniforms [T] float32, final key [2] uint32)."""
    key = np.asarray(rng_key, _U32)
    nrm = np.empty((n_steps, d), np.float32)
    uni = np.empty(n_steps, np.float32)
    for t in range(n_steps):
        ks = split(key, 3)
        key = ks[0]
        nrm[t] = normal(ks[1], (d,))
        uni[t] = uniform(ks[2], ())
    return nrm, uni, key
