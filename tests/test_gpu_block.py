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
    # fp32: both sides sum 5000 residuals in float32, in different orders (the oracle sequentially, the kernel 5 rows per
    # thread of a 1024-thread CTA + a tree); the fp32 oracle itself is 1.4e-3 away from the fp64 one over these 60 steps
    _compare(coll, last, ocoll, olast, tol if prec == "f64" else 1.5 * tol, 1.0 if prec == "f64" else 0.75)
    if prec == "f32":
        # ... and WHY a quarter of the chains may flip: against the float64 oracle every first flip sits inside the band the
        # measured float32 energy error allows (|U| ~ 3.3e3, one ulp = 2.4e-4), and their number is the predicted one
        import flipcheck
        pot = o.make_potential("diamonds", **diamonds_data)
        z0 = ost.z.astype(np.float64)
        o64 = o.arwmh_init(pot, z0)
        orc, _ = flipcheck.oracle_steps(o, o64, pot, nrm, uni, num_warmup=20)
        flipcheck.analyse(_np(coll["accept"]), _np(coll["potential_energy"]), orc, uni, "diamonds block fp32")


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


# ---- correlated Gaussian target: ARWMH (d <= 32) and RAM (d = 200) on the block kernels ----------------
def _random_prec_chol(d, cond, seed):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    ev = np.logspace(0, np.log10(cond), d)
    return np.linalg.cholesky(Q @ np.diag(ev) @ Q.T)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_gaussian_d20_arwmh_shared_draws(prec):
    tdt, ndt, tol = DT[prec]
    d, C, T = 20, 64, 150 if prec == "f64" else 60
    P = _random_prec_chol(d, 100.0, 1)
    sampler = am.ARWMH(models.gaussian, num_chains=C, dtype=tdt)
    state = sampler.init(1, num_warmup=0, init_params=None, model_kwargs=dict(prec_chol=P))
    ost = _oracle_state(state, ndt)
    np.testing.assert_allclose(ost.potential_energy, o.potential_gaussian(ost.z.astype(np.float64), P), rtol=10 * tol)
    rng = np.random.default_rng(2)
    nrm = rng.normal(size=(T, C, d)).astype(ndt)
    uni = rng.random(size=(T, C)).astype(ndt)
    coll, last = sampler.run(state, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)), record_accept=True)
    olast, ocoll = co.arwmh_run(ost, "gaussian", T, draws=(nrm, uni), record_accept=True, prec_chol=P)
    _compare(coll, last, ocoll, olast, tol, 1.0 if prec == "f64" else 0.85)


@pytest.mark.parametrize("kind,d", [("ar1", 200), ("dense", 64)])
def test_ram_gaussian_matches_oracle(kind, d):
    """RAM (not in the reference): GPU vs our float64 oracle under shared draws."""
    P = models.ar1_precision_chol(d, 0.9) if kind == "ar1" else _random_prec_chol(d, 1e3, 3)
    C, T = 8, 40
    rng = np.random.default_rng(5)
    q0 = rng.normal(size=(C, d))
    sampler = am.RAM(models.gaussian, num_chains=C, dtype=torch.float64, init_strategy=am.init_to_value(torch.from_numpy(q0)))
    state = sampler.init(0, num_warmup=0, init_params=None, model_kwargs=dict(prec_chol=P))
    # start from a reasonable scale so that both accepts and rejects occur
    b = am.ChainBatch.from_state(sampler.potential, state)
    b.set_dense_scale(torch.eye(d, dtype=torch.float64) * (2.4 / np.sqrt(d)) * 0.5)
    state = b.to_state()
    nrm = rng.normal(size=(T, C, d))
    uni = rng.random(size=(T, C))
    coll, last = sampler.run(state, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)), record_accept=True)
    pot = o.make_potential("gaussian", prec_chol=P)
    L0 = np.broadcast_to(np.eye(d) * (2.4 / np.sqrt(d)) * 0.5, (C, d, d)).copy()
    zo, Uo, Lo, info = o.ram_run(q0.copy(), pot(q0), L0, pot, T, (nrm, uni), record_accept=True)
    acc_g = _np(coll["accept"])
    assert (acc_g == info["accepts"]).all()
    assert 0.05 < acc_g.mean() < 0.95
    np.testing.assert_allclose(_np(last.z["x"]), zo, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(_np(last.adapt_state.scale), Lo, rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(_np(last.potential_energy), Uo, rtol=1e-8)
    S = _np(last.adapt_state.scale)
    assert np.abs(np.triu(S, 1)).max() == 0 and (np.einsum("cii->ci", S) > 0).all()


def test_ram_adapts_to_target_shape():
    """fp32, Philox: RAM drives the acceptance rate to alpha* = 0.234 from a far-too-wide S = I, the factor stays
    a valid Cholesky factor, and it starts to pick up the target's correlation structure (rank-one learning of
    a 200-dimensional shape is slow: ~0.1 correlation of neighbours after 3e5 steps on this target)."""
    d, C = 200, 64
    P = models.ar1_precision_chol(d, 0.9)
    sampler = am.RAM(models.gaussian, num_chains=C, init_strategy=am.init_to_value(torch.zeros(C, d)))
    state = sampler.init(3, num_warmup=0, init_params=None, model_kwargs=dict(prec_chol=P))
    coll, last = sampler.run(state, 100000, thinning=10000)
    coll2, last2 = sampler.run(last, 5000, collect=(), record_accept=True)
    acc = float(coll2["accept"].float().mean())
    assert 0.18 < acc < 0.30, acc
    S = _np(last2.adapt_state.scale).astype(np.float64)
    assert np.isfinite(S).all() and np.abs(np.triu(S, 1)).max() == 0 and (np.einsum("cii->ci", S) > 0).all()
    SSt = np.einsum("cij,ckj->cik", S, S).mean(0)
    corr1 = np.mean(np.diag(SSt, 1) / np.sqrt(np.diag(SSt)[:-1] * np.diag(SSt)[1:]))
    assert corr1 > 0.03, corr1
    # the chain actually samples N(0, Sigma): pooled marginal variance of the kept draws is O(1)
    x = coll["z"]["x"][-5:].double()
    assert 0.3 < float(x.var()) < 3.0


@pytest.mark.parametrize("cluster", ["2", "4", "8"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_two_cta_cluster_per_chain_matches_the_single_cta_kernel(prec, cluster, monkeypatch):
    """few-chain diamonds: the CTAs of a cluster carry the same chain and split the data rows of the likelihood (arwmh_block.cuh,
    CL = 2, 4, 8).  The only difference to the one-CTA kernel is the order in which the 5000 squared residuals are added: fp64 --
    identical decisions and 1e-9 on the positions over 300 steps; fp32 -- energies of a fixed state agree to 1e-5 relative
    and the chains stay statistically the same (the trajectories themselves part after a first flipped decision)."""
    import adaptive_mcmc_b200 as am
    from adaptive_mcmc_b200 import _lib, models

    tdt = torch.float64 if prec == "f64" else torch.float32
    data = models.synthetic_diamonds(n=5000, k=25, seed=0)
    outs = []
    for cl in ("0", cluster):
        monkeypatch.setenv("AMCMC_BLOCK_CLUSTER", cl)
        s = am.ARWMH(models.diamonds, num_chains=24, dtype=tdt)
        s.impl = _lib.IMPL_BLOCK
        b = s._batch_from_state(s.init(5, num_warmup=100, init_params=None, model_kwargs=data))
        raw = s.run_batch(b, 300, thinning=10, collect_start=5, record_accept=True)
        outs.append((raw["accept"].cpu().numpy(), raw["z"].cpu().numpy(), raw["potential_energy"].cpu().numpy(), b.scale.cpu().numpy(),
                     float(b.macc.mean())))
    a0, z0, u0, sc0, m0 = outs[0]
    a1, z1, u1, sc1, m1 = outs[1]
    if prec == "f64":
        assert (a0 == a1).all()
        np.testing.assert_allclose(z1, z0, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(sc1, sc0, rtol=1e-8, atol=1e-10)
    else:
        np.testing.assert_allclose(u1[0], u0[0], rtol=1e-5)          # first sample: before rounding differences can flip a decision
        same = (a0 == a1).all(axis=0)
        assert same.mean() >= 0.5 and abs(m0 - m1) < 0.03
        # identical decisions, but alpha = exp(U - U') feeds the step-size recursion: the proposals drift apart at the 1e-6 level
        # per step and the far-from-the-mode transient of these 300 steps amplifies it
        err = np.abs(z1[:, :, same] - z0[:, :, same]) / (1 + np.abs(z0[:, :, same]))
        assert err.max() < 5e-2 and np.median(err) < 1e-3, (err.max(), np.median(err))
