"""GPU parity tests for the adaptive stereographic slice sampler (python/kernels/asss.py):
the CUDA kernel through the C ABI vs the NumPy restatement oracle/asss_numpy.py on shared draws, and the
reference's recorded posterior agreement (posteriordb_eight-schools.ipynb cell 29: ASSS matches the ARWMH table)."""
import json
import os

import numpy as np
import pytest
import torch

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models
from oracle import arwmh_numpy as o
from oracle import asss_numpy as oa
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("prec,T,tol", [("f64", 200, 1e-5), ("f32", 60, 1e-3)])
def test_asss_eight_schools_shared_draws(prec, T, tol):
    tdt, ndt = (torch.float64, np.float64) if prec == "f64" else (torch.float32, np.float32)
    C, d = 256, 10
    s = am.ASSS(models.eight_schools, num_chains=C, dtype=tdt)
    st = s.init(5, num_warmup=20, init_params=None)
    assert isinstance(st, am.ASSSState) and st._fields == ("i", "z", "potential_energy", "adapt_state", "as_change", "rng_key")
    q0 = co.init_uniform(5, C, d, dt=ndt)
    pot = o.make_potential("eight_schools")
    ost = oa.asss_init(pot, q0)
    np.testing.assert_allclose(_np(st.potential_energy), ost.potential_energy, rtol=10 * tol)
    rng = np.random.default_rng(11)
    nrm = rng.normal(size=(T, C, d + 1)).astype(ndt)
    uni = rng.random(size=(T, C, 52)).astype(ndt)
    coll, last = s.run(st, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)))
    olast, ocoll = oa.asss_run(ost, pot, T, draws=(nrm, uni), num_warmup=20)
    zg = np.concatenate([_np(v).reshape(T, C, -1) for v in coll["z"].values()], axis=-1)
    err = (np.abs(zg - ocoll["z"]) / (1 + np.abs(ocoll["z"]))).max(axis=(0, 2))
    if prec == "f64":
        assert err.max() < tol
    else:  # a shrinkage comparison pe > t can flip under fp32 round-off; such chains diverge
        assert np.quantile(err, 0.9) < tol, np.quantile(err, [0.5, 0.9, 1.0])
    good = err < 10 * tol
    np.testing.assert_allclose(_np(last.adapt_state.scale)[good], olast.adapt_state.scale[good], rtol=20 * tol, atol=20 * tol)
    np.testing.assert_allclose(_np(last.as_change)[good], olast.as_change[good], rtol=20 * tol, atol=20 * tol)
    np.testing.assert_allclose(_np(last.potential_energy)[good], olast.potential_energy[good], rtol=20 * tol)
    assert int(last.i) == T


def test_asss_philox_stream_and_single_step():
    C, T = 64, 50
    s = am.ASSS(models.eight_schools, num_chains=C, dtype=torch.float64, chain_offset=77)
    st = s.init(2, num_warmup=0, init_params=None)
    pot = o.make_potential("eight_schools")
    ost = oa.asss_init(pot, co.init_uniform(2, C, 10, dt=np.float64, chain_offset=77))
    coll, last = s.run(st, T, thinning=5)
    olast, ocoll = oa.asss_run(ost, pot, T, seed=2, chain_offset=77, thinning=5)
    zg = np.concatenate([_np(v).reshape(T // 5, C, -1) for v in coll["z"].values()], axis=-1)
    err = (np.abs(zg - ocoll["z"]) / (1 + np.abs(ocoll["z"]))).max(axis=(0, 2))
    assert np.quantile(err, 0.9) < 1e-3  # device normals use SFU log/sin/cos
    st2 = s.sample(s.sample(st))  # K = 1 protocol calls are functional
    assert int(st2.i) == 2 and int(st.i) == 0
    c2, l2 = s.run(st, 2, collect=())
    np.testing.assert_allclose(_np(st2.adapt_state.scale), _np(l2.adapt_state.scale), rtol=1e-9, atol=1e-12)


def test_asss_posterior_matches_reference_table():
    tab = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_pins.json")))["eight_schools_arwmh_table"]
    # the reference's ASSS settings: 25k warm-up + 250k samples, thin 25 (run_eight_schools_wasserstein.py:65)
    mcmc = am.MCMC(am.ASSS(models.eight_schools), num_warmup=25000, num_samples=250000, thinning=25, num_chains=128)
    mcmc.run(0, extra_fields=("potential_energy",))
    f = mcmc.get_samples()
    mean = np.array([float(f["mu"].mean())] + [float(v) for v in f["theta_base"].mean(0)])
    std = np.array([float(f["mu"].std())] + [float(v) for v in f["theta_base"].std(0)])
    idx = [0] + list(range(2, 10))
    np.testing.assert_allclose(mean, np.array(tab["mean"])[idx], atol=0.12)
    np.testing.assert_allclose(std, np.array(tab["std"])[idx], atol=0.08)
    g = mcmc.get_samples(group_by_chain=True)
    ess = am.diagnostics.effective_sample_size(g["theta_base"][:32])
    assert float(ess.min()) / 32 > 6000  # reference: n_eff 9275-10281 of 10^4 kept draws
    it = float(mcmc.sampler.mean_shrink_iterations(mcmc.last_state).mean())
    assert 0.2 < it < 5.0, it


# ---- block-per-chain ASSS: diamonds (the reference runs ASSS on diamonds too, run_diamonds_wasserstein.py 'sss') and
# ---- the dense Gaussian family --------------------------------------------------------------------------------------
def _diamonds_start(data, C, rng):
    X, Y = np.asarray(data["X"], np.float64), np.asarray(data["Y"], np.float64)
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    return mode[None] + 0.01 * rng.normal(size=(C, 26))


@pytest.mark.parametrize("prec,T,tol", [("f64", 25, 1e-5), ("f32", 12, 2e-3)])
def test_asss_diamonds_block_shared_draws(prec, T, tol):
    tdt, ndt = (torch.float64, np.float64) if prec == "f64" else (torch.float32, np.float32)
    data = models.synthetic_diamonds(n=5000, k=25, seed=0)
    C, d = 48, 26
    rng = np.random.default_rng(3)
    q0 = _diamonds_start(data, C, rng)
    s = am.ASSS(models.diamonds, num_chains=C, dtype=tdt, init_strategy=am.init_to_value(torch.from_numpy(q0)))
    st = s.init(1, num_warmup=5, init_params=None, model_kwargs=data)
    b = s._batch_from_state(st)
    b.set_dense_scale(torch.eye(d, dtype=tdt) * 0.01)     # scale = I leaves the sphere far too wide for this posterior
    st = s._state_from_batch(b)
    pot = o.make_potential("diamonds", **data)
    z0 = b.z.t().cpu().numpy().astype(ndt)
    ost = oa.asss_init(lambda q: pot(q.astype(np.float64)).astype(ndt), z0)
    ost = ost._replace(adapt_state=ost.adapt_state._replace(scale=np.broadcast_to(np.eye(d, dtype=ndt) * ndt(0.01), (C, d, d)).copy()))
    nrm = rng.normal(size=(T, C, d + 1)).astype(ndt)
    uni = rng.random(size=(T, C, 52)).astype(ndt)
    coll, last = s.run(st, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)))
    olast, ocoll = oa.asss_run(ost, lambda q: pot(q.astype(np.float64)).astype(ndt), T, draws=(nrm, uni), num_warmup=5)
    zg = np.concatenate([_np(v).reshape(T, C, -1) for v in coll["z"].values()], axis=-1)
    err = (np.abs(zg - ocoll["z"]) / (1 + np.abs(ocoll["z"]))).max(axis=(0, 2))
    if prec == "f64":
        assert err.max() < tol, err.max()
    else:
        assert np.quantile(err, 0.8) < tol, np.quantile(err, [0.5, 0.8, 1.0])
    good = err < 10 * tol
    np.testing.assert_allclose(_np(last.adapt_state.scale)[good], olast.adapt_state.scale[good], rtol=50 * tol, atol=50 * tol * 0.01)
    np.testing.assert_allclose(_np(last.adapt_state.loc)[good], olast.adapt_state.loc[good], rtol=20 * tol, atol=20 * tol)
    np.testing.assert_allclose(_np(last.as_change)[good], olast.as_change[good], rtol=50 * tol, atol=1e-6)
    assert int(last.i) == T and (zg[-1] != z0).any()


def test_asss_gaussian_block_philox_and_moments():
    """Dense Gaussian d = 12 on the block kernel: the Philox stream follows the oracle's (same words; the device normals
    use SFU log / sin / cos, so single-precision agreement), and a longer run recovers the target covariance."""
    d, C = 12, 64
    P = models.ar1_precision_chol(d, 0.6)
    s = am.ASSS(models.gaussian, num_chains=C, dtype=torch.float64, chain_offset=5)
    st = s.init(4, num_warmup=0, init_params=None, model_kwargs=dict(prec_chol=P))
    Pn = np.asarray(P, np.float64)
    pot = lambda q: 0.5 * ((q @ np.tril(Pn)) ** 2).sum(1)
    ost = oa.asss_init(pot, co.init_uniform(4, C, d, dt=np.float64, chain_offset=5))
    np.testing.assert_allclose(_np(st.potential_energy), ost.potential_energy, rtol=1e-10)
    coll, last = s.run(st, 30, thinning=3)
    olast, ocoll = oa.asss_run(ost, pot, 30, seed=4, chain_offset=5, thinning=3)
    err = (np.abs(_np(coll["z"]["x"]) - ocoll["z"]) / (1 + np.abs(ocoll["z"]))).max(axis=(0, 2))
    assert np.quantile(err, 0.9) < 1e-3, np.quantile(err, [0.5, 0.9, 1.0])  # device normals use SFU log/sin/cos
    good = err < 1e-3
    np.testing.assert_allclose(_np(last.adapt_state.scale)[good], olast.adapt_state.scale[good], rtol=2e-2, atol=2e-3)
    # moments
    # (while the Robbins-Monro weights are still large the sample is over-dispersed -- 1.36 on the diagonal after
    # 2k + 4k steps, in fp64 as in fp32 -- so the moments are checked on a run of the reference's length)
    s2 = am.ASSS(models.gaussian, num_chains=512, dtype=torch.float32)
    st2 = s2.init(9, num_warmup=20000, init_params=None, model_kwargs=dict(prec_chol=P))
    coll2, _ = s2.run(st2, 60000, thinning=20, collect_start=20000)
    x = coll2["z"]["x"].double().reshape(-1, d).cpu().numpy()
    cov = np.linalg.inv(np.tril(Pn) @ np.tril(Pn).T)
    emp = np.cov(x.T)
    assert np.abs(emp - cov).max() < 0.04, np.abs(emp - cov).max()


def test_asss_logscale_collection_and_reference_pickle(tmp_path):
    """collect_states_logscale (python/utils/kernel_utils.py:20-38) with the slice sampler, written in the reference's
    pickle layout: `kernels.asss.ASSSState` records with [sample, ...] NumPy leaves."""
    import pickle
    import sys
    import types
    from collections import namedtuple

    from adaptive_mcmc_b200.utils import io as amio

    s = am.ASSS(models.eight_schools, lr_decay=0.5, num_chains=6)
    states = am.collect_states_logscale(3, s, dict(y=models.eight_schools.Y, sigma=models.eight_schools.SIGMA), n_pow=3)
    grid = am.ns_logscale(3)
    assert type(states).__name__ == "ASSSState" and states.as_change.shape == (len(grid), 6)
    np.testing.assert_array_equal(_np(states.i).ravel(), grid.numpy())
    assert states.adapt_state.scale.shape == (len(grid), 6, 10, 10) and states.z["theta_base"].shape == (len(grid), 6, 8)
    path = tmp_path / "run4.pkl"
    amio.save_states(states, str(path), chain=4)
    # load it the way the reference environment would: with ITS record classes importable as kernels.asss.*
    mod_k, mod_a = types.ModuleType("kernels"), types.ModuleType("kernels.asss")
    mod_a.ASSSState = namedtuple("ASSSState", ["i", "z", "potential_energy", "adapt_state", "as_change", "rng_key"])
    mod_a.ASSSAdaptState = namedtuple("ASSSAdaptState", ["loc", "scale"])
    mod_a.ASSSState.__module__ = mod_a.ASSSAdaptState.__module__ = "kernels.asss"
    sys.modules["kernels"], sys.modules["kernels.asss"] = mod_k, mod_a
    try:
        got = pickle.load(open(path, "rb"))
    finally:
        sys.modules.pop("kernels.asss"), sys.modules.pop("kernels")
    assert isinstance(got, mod_a.ASSSState) and isinstance(got.adapt_state, mod_a.ASSSAdaptState)
    np.testing.assert_allclose(got.potential_energy, _np(states.potential_energy)[:, 4])
    assert got.adapt_state.scale.shape == (len(grid), 10, 10) and got.z["mu"].shape == (len(grid),)


def test_asss_sample_Pnx_frozen_kernel():
    """ASSS.sample_Pnx (asss.py:279-315): the slice sampler with a FROZEN (loc, scale).  (i) exact N(0, I_2) draws pushed
    through P^n stay N(0, I_2) (asumptions_check.ipynb cells 32-33 run the ASSS twin of the ARWMH check); (ii) the frozen
    step equals the oracle's on shared draws and leaves the adaptation state untouched, on both kernel families."""
    from scipy import stats

    pot = models.std_normal.bind(d=2, dtype=torch.float32)
    s = am.ASSS(potential_fn=pot)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(100000, 2, generator=g)
    adapt = am.ASSSAdaptState(torch.tensor([0.3, -0.2]), torch.tensor([[1.2, 0.0], [0.4, 0.8]]))
    out = s.sample_Pnx(5, x, adapt, n=3, n_samples=2)
    assert out.shape == (100000, 2, 2)
    y = out.reshape(-1, 2).double().cpu().numpy()
    assert np.abs(y.mean(0)).max() < 0.01 and np.abs(y.std(0) - 1).max() < 0.01 and abs(np.corrcoef(y.T)[0, 1]) < 0.01
    assert stats.kstest(y[::4, 0], "norm").pvalue > 1e-3 and stats.kstest(y[::4, 1], "norm").pvalue > 1e-3
    assert "Potential Energy" in s.get_diagnostics_str(s.init(0, 0, init_params=torch.zeros(4, 2)))
    # (ii) shared draws, frozen, register kernel (eight_schools) and block kernel (Gaussian d = 12)
    rng = np.random.default_rng(2)
    for family, kw, name, d in ((models.eight_schools, {}, "eight_schools", 10),
                                (models.gaussian, dict(prec_chol=models.ar1_precision_chol(12, 0.5)), None, 12)):
        C, T = 64, 15
        smp = am.ASSS(family, num_chains=C, dtype=torch.float64)
        st = smp.init(3, num_warmup=0, init_params=None, model_kwargs=kw)
        b = smp._batch_from_state(st)
        scale0 = torch.tril(torch.from_numpy(rng.normal(size=(d, d)) * 0.1 + np.eye(d)))
        b.set_dense_scale(scale0)
        loc0, sc0 = b.loc.clone(), b.scale.clone()
        nrm, uni = rng.normal(size=(T, C, d + 1)), rng.random(size=(T, C, 52))
        dr = smp._draws_to_device_layout((torch.from_numpy(nrm), torch.from_numpy(uni)))
        z0 = b.z.t().cpu().numpy().copy()
        smp.run_batch(b, T, collect=(), draws=dr, adapt=False)
        assert torch.equal(b.loc, loc0) and torch.equal(b.scale, sc0)
        if name:
            potf = o.make_potential(name)
        else:
            Pn = np.tril(np.asarray(kw["prec_chol"], np.float64))
            potf = lambda q: 0.5 * ((q @ Pn) ** 2).sum(1)
        ost = oa.asss_init(potf, z0)
        ost = ost._replace(adapt_state=oa.ASSSAdaptState(z0.copy(), np.broadcast_to(scale0.numpy(), (C, d, d)).copy()))
        olast, _ = oa.asss_run(ost, potf, T, draws=(nrm, uni), adapt=False)
        np.testing.assert_allclose(b.z.t().cpu().numpy(), olast.z, rtol=1e-6, atol=1e-8)


def test_asss_diamonds_through_the_mcmc_driver():
    """run_diamonds_wasserstein.py 'sss' in miniature: MCMC(ASSS(model)).run(key, **data) with the diamonds model."""
    data = models.synthetic_diamonds(n=5000, k=25, seed=0)
    rng = np.random.default_rng(0)
    q0 = _diamonds_start(data, 32, rng)
    mcmc = am.MCMC(am.ASSS(models.diamonds, init_strategy=am.init_to_value(torch.from_numpy(q0))), num_warmup=1500, num_samples=1500,
                   thinning=15, num_chains=32)
    mcmc.run(1, **data, extra_fields=("potential_energy",))
    smp = mcmc.get_samples(group_by_chain=True)
    assert smp["b"].shape == (32, 100, 24) and smp["sigma"].shape == (32, 100)
    sig = float(smp["sigma"][:, 50:].mean())
    assert abs(sig - 0.123) < 0.01, sig                      # the data were generated with sigma = 0.123
    pe = mcmc.get_extra_fields(group_by_chain=True)["potential_energy"]
    assert torch.isfinite(pe).all() and float(pe[:, -1].median()) < -3.2e3    # at the mode U ~ -3.28e3
