"""GPU parity tests for the block-per-chain kernel (diamonds on CUDA cores: the exact path).
Same criteria as tests/test_gpu_small.py; the oracle is the C restatement of the reference."""
import numpy as np
import pytest
import torch

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models
from oracle import arwmh_numpy as o
from oracle import c_oracle as co
from test_gpu_small import DT, _compare, _np, _oracle_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def diamonds_data():
    return models.synthetic_diamonds(n=5000, k=25, seed=0)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_diamonds_potential(prec, diamonds_data):
    tdt, ndt, tol = DT[prec]
    pot = models.diamonds.bind(dtype=tdt, **diamonds_data)
    assert pot.dim == 26 and list(pot.sites) == ["Intercept", "b", "sigma"]
    rng = np.random.default_rng(0)
    q = rng.normal(size=(64, 26)) * 0.2
    q[:, 0] += 7.8
    q[:, -1] = -2.0 + 0.1 * rng.normal(size=64)
    ref = o.potential_diamonds(q, diamonds_data["X"], diamonds_data["Y"])
    got = _np(pot(torch.from_numpy(q)))
    np.testing.assert_allclose(got, ref, rtol=1e-11 if prec == "f64" else 2e-5)
    # far from the mode (the init region) too
    q2 = rng.uniform(-2, 2, size=(64, 26))
    np.testing.assert_allclose(_np(pot(torch.from_numpy(q2))), o.potential_diamonds(q2, diamonds_data["X"], diamonds_data["Y"]),
                               rtol=1e-11 if prec == "f64" else 2e-5)


@pytest.mark.parametrize("prec,T", [("f64", 200), ("f32", 60)])
def test_diamonds_shared_draws(prec, T, diamonds_data):
    tdt, ndt, tol = DT[prec]
    C, d = 64, 26  # BASELINE.json configs[1]: 64 chains
    if prec == "f64":
        sampler = am.ARWMH(models.diamonds, num_chains=C, dtype=tdt)
    else:
        # fp32 starts near the mode: from U(-2,2) the potential is ~1e4-1e6 where fp32 resolves U to only
        # ~1e-2, and the fp32 ORACLE itself already deviates from the fp64 oracle by 1.4e-3 over 60 steps
        X, Y = diamonds_data["X"], diamonds_data["Y"]
        Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
        ols = np.linalg.lstsq(Xc, Y, rcond=None)[0]
        q0 = np.concatenate([ols, [np.log(0.123)]])[None] + 0.01 * np.random.default_rng(5).normal(size=(C, d))
        sampler = am.ARWMH(models.diamonds, num_chains=C, dtype=tdt, init_strategy=am.init_to_value(torch.from_numpy(q0)))
    state = sampler.init(11, num_warmup=20, init_params=None, model_kwargs=diamonds_data)
    ost = _oracle_state(state, ndt)
    if prec == "f64":
        np.testing.assert_array_equal(ost.z, co.init_uniform(11, C, d, dt=ndt))
    np.testing.assert_allclose(ost.potential_energy,
                               o.potential_diamonds(ost.z.astype(np.float64), diamonds_data["X"], diamonds_data["Y"]),
                               rtol=1e-10 if prec == "f64" else 1e-4)
    rng = np.random.default_rng(3)
    nrm = rng.normal(size=(T, C, d)).astype(ndt)
    uni = rng.random(size=(T, C)).astype(ndt)
    coll, last = sampler.run(state, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)), record_accept=True)
    olast, ocoll = co.arwmh_run(ost, "diamonds", T, draws=(nrm, uni), record_accept=True, num_warmup=20, **diamonds_data)
    _compare(coll, last, ocoll, olast, tol, 1.0 if prec == "f64" else 0.75)


def test_diamonds_philox_segmentation(diamonds_data):
    C = 16
    sampler = am.ARWMH(models.diamonds, num_chains=C, dtype=torch.float64)
    s0 = sampler.init(2, num_warmup=0, init_params=None, model_kwargs=diamonds_data)
    ost = _oracle_state(s0, np.float64)
    coll, full = sampler.run(s0, 90, thinning=9, record_accept=True)
    olast, ocoll = co.arwmh_run(ost, "diamonds", 90, seed=2, thinning=9, record_accept=True, **diamonds_data)
    _compare(coll, full, ocoll, olast, 1e-3, 0.9)  # device normals use SFU log/sin/cos
    s = s0
    for k in (1, 29, 60):
        s = sampler.run(s, k, collect=())[1]
    np.testing.assert_allclose(_np(s.adapt_state.scale), _np(full.adapt_state.scale), rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(_np(s.as_change), _np(full.as_change), rtol=1e-7)


def test_diamonds_sampler_converges(diamonds_data):
    # 64 chains (BASELINE.json configs[1]) started at the mode region; after adaptation the draws must
    # match the analytic conditional-Gaussian posterior of the regression coefficients
    X, Y = diamonds_data["X"], diamonds_data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    ols = np.linalg.lstsq(Xc, Y, rcond=None)[0]
    C, d = 64, 26
    q0 = np.concatenate([ols, [np.log(0.123)]])[None] + 0.001 * np.random.default_rng(0).normal(size=(C, d))
    sampler = am.ARWMH(models.diamonds, init_strategy=am.init_to_value(torch.from_numpy(q0)))
    mcmc = am.MCMC(sampler, num_warmup=200000, num_samples=100000, thinning=100, num_chains=C)
    mcmc.run(0, **diamonds_data, extra_fields=("potential_energy",))
    s = mcmc.get_samples()
    sig = float(s["sigma"].mean())
    assert abs(sig - 0.123) < 0.004
    G = Xc[:, 1:].T @ Xc[:, 1:]
    post_cov = sig**2 * np.linalg.inv(G + sig**2 * np.eye(24))
    ridge = np.linalg.solve(G + sig**2 * np.eye(24), Xc[:, 1:].T @ (Y - Y.mean()))
    sd = np.sqrt(np.diag(post_cov))
    b_mean = s["b"].double().mean(0).cpu().numpy()
    b_sd = s["b"].double().std(0).cpu().numpy()
    assert (np.abs(b_mean - ridge) / sd).max() < 1.0, (np.abs(b_mean - ridge) / sd)
    assert 0.6 < (b_sd / sd).min() and (b_sd / sd).max() < 1.5, b_sd / sd
    assert abs(float(s["Intercept"].mean()) - Y.mean()) < 0.002
    acc = float(mcmc.last_state.mean_accept_prob.mean())
    assert 0.15 < acc < 0.35
