"""The reference's random stream on the GPU (csrc/jax_rng.cu, `amcmc_jax_draws`; `ARWMH(..., rng="jax")`) against
oracle/jax_random.py -- the NumPy restatement of jax.random's threefry2x32 stream that tests/test_jax_random.py pins to the
Random123 vectors and to the values the JAX documentation prints.  Then whole trajectories: a chain started from
`PRNGKey(seed)` on the GPU follows the oracle fed with the same key's draws (python/kernels/arwmh.py:162-165,174)."""
import ctypes as C

import numpy as np
import pytest
import torch

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import _lib, models
from adaptive_mcmc_b200.utils import jax_prng
from oracle import arwmh_numpy as o
from oracle import jax_random as jr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,T", [(10, 40), (1, 25), (7, 30), (26, 12)])
def test_draws_match_the_restated_jax_stream(d, T):
    seeds = [0, 1, 42, 2**31 + 5, 123456789]
    keys = np.stack([jr.prng_key(s) for s in seeds])                       # [C, 2]
    Cn = len(seeds)
    dk = torch.from_numpy(np.ascontiguousarray(keys.T).view(np.int32)).cuda()  # [2, C]
    nrm = torch.empty(T, d, Cn, device="cuda")
    uni = torch.empty(T, Cn, device="cuda")
    _lib.check(_lib.lib().amcmc_jax_draws(dk.data_ptr(), Cn, d, T, _lib.AMCMC_F32, nrm.data_ptr(), uni.data_ptr(),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)), "amcmc_jax_draws")
    kout = dk.cpu().numpy().view(np.uint32).T
    for c, s in enumerate(seeds):
        n_o, u_o, k_o = jr.arwmh_draws(jr.prng_key(s), d, T)
        np.testing.assert_array_equal(uni[:, c].cpu().numpy(), u_o)                      # pure bit manipulation: exact
        np.testing.assert_allclose(nrm[:, :, c].cpu().numpy(), n_o, rtol=0, atol=6e-7)   # erfinv: log1pf of libm vs CUDA
        np.testing.assert_array_equal(kout[c], k_o)                                       # the carried key
    assert jax_prng.split(jax_prng.prng_key(7), 5).tolist() == jr.split(jr.prng_key(7), 5).tolist()


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_chains_started_from_a_jax_key_follow_the_oracle(prec):
    tdt, ndt, tol = (torch.float64, np.float64, 1e-5) if prec == "f64" else (torch.float32, np.float32, 1e-3)
    Cn, T, d = 64, 150, 10
    q0 = np.random.default_rng(1).uniform(-2, 2, size=(Cn, d))
    s = am.ARWMH(models.eight_schools, num_chains=Cn, dtype=tdt, rng="jax", init_strategy=am.init_to_value(torch.from_numpy(q0)))
    st = s.init(jr.prng_key(11), num_warmup=30, init_params=None)
    keys = jax_prng.chain_keys(jr.prng_key(11), Cn)                      # NumPyro: one key per chain from split(rng_key, C)
    np.testing.assert_array_equal(st.rng_key.cpu().numpy(), keys.astype(np.int64))
    coll, last = s.run(st, T, record_accept=True)
    nrm = np.stack([jr.arwmh_draws(keys[c], d, T)[0] for c in range(Cn)], axis=1)
    uni = np.stack([jr.arwmh_draws(keys[c], d, T)[1] for c in range(Cn)], axis=1)
    ost = o.arwmh_init(o.make_potential("eight_schools"), q0.astype(ndt))
    olast, ocoll = o.arwmh_run(ost, o.make_potential("eight_schools"), T, draws=(nrm.astype(ndt), uni.astype(ndt)),
                               record_accept=True, num_warmup=30)
    same = (coll["accept"].cpu().numpy() == ocoll["accepts"]).all(axis=0)
    assert same.mean() >= (1.0 if prec == "f64" else 0.9), same.mean()
    zg = np.concatenate([v.cpu().numpy().reshape(T, Cn, -1) for v in coll["z"].values()], axis=-1)
    err = (np.abs(zg - ocoll["z"]) / (1 + np.abs(ocoll["z"]))).max(axis=(0, 2))[same]
    assert err.max() < tol, err.max()
    # the carried keys are the oracle's, and a second call continues the stream (segmentation invariance)
    k_end = np.stack([jr.arwmh_draws(keys[c], d, T)[2] for c in range(Cn)])
    np.testing.assert_array_equal(last.rng_key.cpu().numpy(), k_end.astype(np.int64))
    st2 = s.init(jr.prng_key(11), num_warmup=30, init_params=None)
    _, mid = s.run(st2, 70, collect=())
    _, end = s.run(mid, T - 70, collect=())
    np.testing.assert_array_equal(end.rng_key.cpu().numpy(), last.rng_key.cpu().numpy())
    if prec == "f64":
        np.testing.assert_allclose(end.adapt_state.loc.cpu().numpy(), last.adapt_state.loc.cpu().numpy(), rtol=1e-9, atol=1e-10)


def test_jax_stream_drives_the_tensor_core_kernel_too():
    """rng='jax' goes through the external-draws mode, so every kernel family accepts it -- here the tcgen05 diamonds kernel."""
    data = models.synthetic_diamonds(n=5000, k=25, seed=0)
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    Cn = 512
    q0 = mode[None] + 0.004 * np.random.default_rng(2).normal(size=(Cn, 26))
    res = {}
    for impl in (_lib.IMPL_TENSOR, _lib.IMPL_BLOCK):
        s = am.ARWMH(models.diamonds, num_chains=Cn, rng="jax", init_strategy=am.init_to_value(torch.from_numpy(q0)))
        s.impl = impl
        st = s.init(3, num_warmup=0, init_params=None, model_kwargs=data)
        b = am.ChainBatch.from_state(s.potential, st)
        b.set_dense_scale(torch.eye(26) * 0.002)
        raw = s.run_batch(b, 10, collect=(), record_accept=True)
        res[impl] = (raw["accept"].cpu().numpy(), b.z.clone(), b.jax_keys.clone())
    same = (res[_lib.IMPL_TENSOR][0] == res[_lib.IMPL_BLOCK][0]).all(axis=0)
    assert same.mean() > 0.95 and torch.equal(res[_lib.IMPL_TENSOR][2], res[_lib.IMPL_BLOCK][2])
