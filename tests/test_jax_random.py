"""The reference's random stream (jax.random, threefry2x32) restated in oracle/jax_random.py, pinned by

* the Random123 known-answer vectors for Threefry-2x32-20 (the ones JAX's own tests/random_test.py::testThreefry2x32 uses);
* the values the JAX documentation prints for its default (non-partitionable threefry) PRNG:
  PRNGKey(0): split -> [4146024105 967050713] [2718843009 1272950319]; normal(key, (1,)) -> [-0.20584226];
  uniform(key) -> 0.41845703;  PRNGKey(42): normal(key) -> -0.18471177; split -> [2465931498 3679230171]
  [255383827 267815257]; normal(subkey) -> 1.3694694; normal(key, (3,)) -> [0.18693547 -1.2806505 -1.5593132].
  (Recalled from the docs -- there is no JAX in this image -- but they only come out right if the block function, the
  counter pairing of `threefry_2x32`, `split`, `random_bits`, the mantissa trick of `uniform` and XLA's float32 erfinv
  polynomial are all right; scripts/jax_bridge.py re-checks against the real library when it is available.)
"""
import numpy as np

from oracle import jax_random as jr


def test_threefry2x32_random123_known_answers():
    kat = [((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, want in kat:
        y0, y1 = jr.threefry2x32(key, [ctr[0]], [ctr[1]])
        assert (int(y0[0]), int(y1[0])) == want


def test_jax_documented_values():
    k0 = jr.prng_key(0)
    assert k0.tolist() == [0, 0]
    assert jr.split(k0, 2).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    np.testing.assert_allclose(jr.normal(k0, (1,)), [-0.20584226], rtol=0, atol=2e-8)
    np.testing.assert_allclose(jr.uniform(k0), 0.41845703, rtol=0, atol=1e-8)
    k42 = jr.prng_key(42)
    np.testing.assert_allclose(jr.normal(k42), -0.18471177, rtol=0, atol=2e-8)
    ks = jr.split(k42, 2)
    assert ks.tolist() == [[2465931498, 3679230171], [255383827, 267815257]]
    np.testing.assert_allclose(jr.normal(ks[1]), 1.3694694, rtol=0, atol=2e-7)
    np.testing.assert_allclose(jr.normal(k42, (3,)), [0.18693547, -1.2806505, -1.5593132], rtol=0, atol=2e-7)


def test_stream_properties_and_arwmh_draws():
    # odd counter lengths are padded with one zero and the last output dropped: a prefix property does NOT hold,
    # the pairing is (first half, second half) -- pin both facts
    k = jr.prng_key(7)
    a5, a6 = jr.random_bits(k, 5), jr.random_bits(k, 6)
    assert a5.shape == (5,) and not np.array_equal(a5, a6[:5])
    y0, y1 = jr.threefry2x32(k, np.array([0, 1, 2], np.uint32), np.array([3, 4, 5], np.uint32))
    assert np.array_equal(a6, np.concatenate([y0, y1]))
    # uniform in [0, 1), normal moments over a long stream
    u = jr.uniform(k, (200000,))
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 3e-3
    z = jr.normal(k, (200000,)).astype(np.float64)
    assert abs(z.mean()) < 8e-3 and abs(z.std() - 1.0) < 5e-3 and np.isfinite(z).all()
    from scipy import special

    np.testing.assert_allclose(jr.erfinv_f32(np.linspace(-0.999, 0.999, 101)), special.erfinv(np.linspace(-0.999, 0.999, 101)),
                               rtol=3e-6, atol=1e-7)
    # the per-step key chain of ARWMH.sample (arwmh.py:162): key <- split(key, 3)[0]
    nrm, uni, key = jr.arwmh_draws(jr.prng_key(3), 10, 4)
    ks = jr.split(jr.prng_key(3), 3)
    np.testing.assert_array_equal(nrm[0], jr.normal(ks[1], (10,)))
    assert uni[0] == jr.uniform(ks[2])
    k1 = jr.split(ks[0], 3)
    np.testing.assert_array_equal(nrm[1], jr.normal(k1[1], (10,)))
    assert nrm.shape == (4, 10) and uni.shape == (4,) and key.shape == (2,)
