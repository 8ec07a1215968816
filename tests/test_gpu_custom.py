"""Custom model families (adaptive_mcmc_b200/custom.py): a potential written as a CUDA device function, compiled into a
plugin with the fused thread-per-chain kernels, against the NumPy restatement of the reference step run on the same
potential written in NumPy -- the reference accepts arbitrary model functions (arwmh.py:43-78), this is its counterpart."""
import os

import numpy as np
import pytest
import torch

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import custom, models
from oracle import arwmh_numpy as o
from oracle import asss_numpy as oa

LOGISTIC_SRC = '''
    R u = (R)0.5 * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) * (R)0.04;   // beta ~ N(0, 5^2)
    for (int64_t i = 0; i < n1; ++i) {
      const R eta = q[0] + q[1] * a0[2 * i] + q[2] * a0[2 * i + 1];
      const R sp = eta > (R)0 ? eta + Num<R>::log1p(Num<R>::exp(-eta)) : Num<R>::log1p(Num<R>::exp(eta));   // softplus
      u += sp - a1[i] * eta;                                                // -log Bernoulli(y_i | sigmoid(eta))
    }
    return u;'''

EIGHT_SCHOOLS_SRC = '''
    // q = [mu, t = log tau, eta_1..8]; a0 = y, a1 = sigma   (run_eight_schools_lr_decay.py:26-35, non-centred)
    const R mu = q[0], t = q[1], tau = Num<R>::exp(t);
    R u = (R)0.5 * mu * mu * (R)0.04 + Num<R>::log1p(tau * tau * (R)0.04) - t;
    for (int j = 0; j < 8; ++j) {
      const R r = (a0[j] - mu - tau * q[2 + j]) / a1[j];
      u += (R)0.5 * (q[2 + j] * q[2 + j] + r * r);
    }
    return u;'''


def _logistic_data(n=200, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, 2))
    p = 1 / (1 + np.exp(-(0.5 + 1.2 * x[:, 0] - 0.7 * x[:, 1])))
    y = (rng.random(n) < p).astype(np.float64)
    return x, y


def _logistic_potential(x, y):
    def pot(q):
        eta = q[:, :1] + q[:, 1:2] * x[None, :, 0] + q[:, 2:3] * x[None, :, 1]
        return 0.5 * (q**2).sum(1) / 25 + (np.logaddexp(0, eta) - y[None] * eta).sum(1)
    return pot


def test_plugin_builds_without_a_gpu_and_rejects_bad_source(built):
    so = custom.build_plugin("t_quadratic", 2, "    return (R)0.5 * (q[0] * q[0] + q[1] * q[1]);")
    import ctypes
    L = ctypes.CDLL(so)
    assert L.amcmc_plugin_dim() == 2
    for sym in ("amcmc_plugin_run", "amcmc_plugin_init", "amcmc_plugin_potential", "amcmc_plugin_abi"):
        getattr(L, sym)
    assert custom.build_plugin("t_quadratic", 2, "    return (R)0.5 * (q[0] * q[0] + q[1] * q[1]);") == so  # cached
    with pytest.raises(ValueError, match="nvcc failed"):
        custom.build_plugin("t_broken", 2, "    return undeclared_symbol;")
    with pytest.raises(ValueError):
        am.custom_model("bad name", [("x", (2,))], "return 0;")
    with pytest.raises(ValueError):
        am.custom_model("toolarge", [("x", (40,))], "return 0;")


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_custom_logistic_matches_oracle(prec):
    tdt, ndt, tol = (torch.float64, np.float64, 1e-8) if prec == "f64" else (torch.float32, np.float32, 2e-3)
    x, y = _logistic_data()
    fam = am.custom_model("t_logistic3", [("beta", (3,))], LOGISTIC_SRC, arrays=["x", "y"])
    C, T, d = 256, 60, 3
    s = am.ARWMH(fam, num_chains=C, dtype=tdt)
    st = s.init(3, num_warmup=10, init_params=None, model_kwargs=dict(x=x, y=y))
    pot = _logistic_potential(x, y)
    q0 = st.z["beta"].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(st.potential_energy.cpu().numpy(), pot(q0), rtol=1e-12 if prec == "f64" else 2e-5)
    np.testing.assert_allclose(s.potential(torch.from_numpy(q0)).cpu().numpy(), pot(q0), rtol=1e-12 if prec == "f64" else 2e-5)
    rng = np.random.default_rng(5)
    nrm, uni = rng.normal(size=(T, C, d)).astype(ndt), rng.random(size=(T, C)).astype(ndt)
    coll, last = s.run(st, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)), record_accept=True)
    ost = o.arwmh_init(lambda q: pot(q.astype(np.float64)).astype(ndt), q0.astype(ndt))
    olast, ocoll = o.arwmh_run(ost, lambda q: pot(q.astype(np.float64)).astype(ndt), T, draws=(nrm, uni), record_accept=True, num_warmup=10)
    same = (coll["accept"].cpu().numpy().astype(bool) == ocoll["accepts"].astype(bool)).all(axis=0)
    assert same.mean() > (0.999 if prec == "f64" else 0.9)
    for g, r in ((last.z["beta"], olast.z), (last.adapt_state.loc, olast.adapt_state.loc), (last.adapt_state.scale, olast.adapt_state.scale),
                 (last.adapt_state.log_step_size, olast.adapt_state.log_step_size), (last.mean_accept_prob, olast.mean_accept_prob)):
        g, r = g.cpu().numpy()[same], np.asarray(r)[same]
        err = (np.abs(g - r) / (1 + np.abs(r))).reshape(g.shape[0], -1).max(1)
        assert np.quantile(err, 0.99) < tol and err.max() < 20 * tol, (err.max(), np.quantile(err, 0.99))


@pytest.mark.gpu
def test_custom_eight_schools_equals_builtin_family():
    """The same posterior written as a plugin: identical potential (up to the folded constant) and, with shared draws,
    the same chains as the built-in eight_schools family."""
    fam = am.custom_model("t_eight_schools", [("mu", ()), ("tau", ()), ("theta_base", (8,))], EIGHT_SCHOOLS_SRC, arrays=["y", "sigma"],
                          postprocess=lambda z, data: models.eight_schools.postprocess(z, data))
    data = dict(y=models.eight_schools.Y, sigma=models.eight_schools.SIGMA)
    C, T, d = 512, 80, 10
    rng = np.random.default_rng(1)
    q0 = rng.uniform(-2, 2, size=(C, d))
    nrm, uni = rng.normal(size=(T, C, d)), rng.random(size=(T, C))
    out = {}
    for name, model in (("plugin", fam), ("builtin", models.eight_schools)):
        s = am.ARWMH(model, num_chains=C, dtype=torch.float64, init_strategy=am.init_to_value(torch.from_numpy(q0)))
        st = s.init(0, num_warmup=0, init_params=None, model_kwargs=data)
        coll, last = s.run(st, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)), record_accept=True)
        out[name] = (st.potential_energy.cpu().numpy(), coll["accept"].cpu().numpy(), last)
    shift = out["builtin"][0] - out["plugin"][0]
    assert np.abs(shift - shift[0]).max() < 1e-10          # the potentials differ by the constants the plugin leaves out
    np.testing.assert_array_equal(out["plugin"][1], out["builtin"][1])
    for site in ("mu", "tau", "theta_base"):
        np.testing.assert_allclose(out["plugin"][2].z[site].cpu().numpy(), out["builtin"][2].z[site].cpu().numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(out["plugin"][2].adapt_state.scale.cpu().numpy(), out["builtin"][2].adapt_state.scale.cpu().numpy(), rtol=1e-8, atol=1e-10)
    # driver, postprocess and the second sampler on the plugin family
    means = {}
    for name, model in (("plugin", fam), ("builtin", models.eight_schools)):
        mcmc = am.MCMC(am.ASSS(model, dtype=torch.float64), num_warmup=5000, num_samples=20000, thinning=20, num_chains=256)
        mcmc.run(2, **data)
        smp = mcmc.get_samples()
        assert smp["theta"].shape == (256 * 1000, 8) and torch.isfinite(smp["tau"]).all()
        means[name] = (float(smp["mu"].mean()), float(smp["tau"].mean()), float(smp["theta"].mean()))
    # same Philox streams and the same potential up to a constant: the two runs are the same chains
    np.testing.assert_allclose(means["plugin"], means["builtin"], rtol=1e-6)
    assert abs(means["plugin"][0] - 4.40) < 0.2 and abs(means["plugin"][1] - 3.60) < 0.25


@pytest.mark.gpu
def test_custom_model_asss_matches_oracle():
    x, y = _logistic_data(120, seed=4)
    fam = am.custom_model("t_logistic3", [("beta", (3,))], LOGISTIC_SRC, arrays=["x", "y"])
    C, T, d = 128, 25, 3
    s = am.ASSS(fam, num_chains=C, dtype=torch.float64)
    st = s.init(8, num_warmup=0, init_params=None, model_kwargs=dict(x=x, y=y))
    pot = _logistic_potential(x, y)
    rng = np.random.default_rng(6)
    nrm, uni = rng.normal(size=(T, C, d + 1)), rng.random(size=(T, C, 52))
    coll, last = s.run(st, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)))
    q0 = st.z["beta"].cpu().numpy()
    ost = oa.asss_init(pot, q0)
    olast, _ = oa.asss_run(ost, pot, T, draws=(nrm, uni))
    np.testing.assert_allclose(last.z["beta"].cpu().numpy(), olast.z, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(last.adapt_state.scale.cpu().numpy(), olast.adapt_state.scale, rtol=1e-6, atol=1e-8)


LOGISTIC_ROW = '''
    const R eta = q[0] + q[1] * a0[2 * i] + q[2] * a0[2 * i + 1];
    const R sp = eta > (R)0 ? eta + Num<R>::log1p(Num<R>::exp(-eta)) : Num<R>::log1p(Num<R>::exp(eta));
    return sp - a1[i] * eta;'''
LOGISTIC_PRIOR = "    return (R)0.02 * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);"


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["arwmh", "asss"])
def test_custom_block_model_matches_oracle(kind):
    """A data-heavy custom model (logistic regression, 20,000 rows) on the CTA-per-chain kernels: the rows of the
    likelihood are spread over the chain's CTA like the diamonds likelihood."""
    x, y = _logistic_data(20000, seed=2)
    fam = am.custom_model("t_logistic_rows", [("beta", (3,))], arrays=["x", "y"], rows="y", row_term=LOGISTIC_ROW, prior=LOGISTIC_PRIOR)
    pot = _logistic_potential(x, y)
    C, d = 96, 3
    rng = np.random.default_rng(8)
    q0 = np.array([0.5, 1.2, -0.7])[None] + 0.05 * rng.normal(size=(C, d))
    cls = am.ARWMH if kind == "arwmh" else am.ASSS
    s = cls(fam, num_chains=C, dtype=torch.float64, init_strategy=am.init_to_value(torch.from_numpy(q0)))
    st = s.init(0, num_warmup=5, init_params=None, model_kwargs=dict(x=x, y=y))
    np.testing.assert_allclose(st.potential_energy.cpu().numpy(), pot(q0), rtol=1e-11)
    b = s._batch_from_state(st)
    b.set_dense_scale(torch.eye(d, dtype=torch.float64) * 0.02)
    st = s._state_from_batch(b)
    T = 30
    if kind == "arwmh":
        nrm, uni = rng.normal(size=(T, C, d)), rng.random(size=(T, C))
        coll, last = s.run(st, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)), record_accept=True)
        ost = o.arwmh_init(pot, q0)
        ost = ost._replace(adapt_state=ost.adapt_state._replace(scale=np.broadcast_to(np.eye(d) * 0.02, (C, d, d)).copy()))
        olast, ocoll = o.arwmh_run(ost, pot, T, draws=(nrm, uni), record_accept=True, num_warmup=5)
        np.testing.assert_array_equal(coll["accept"].cpu().numpy().astype(bool), ocoll["accepts"].astype(bool))
        assert 0.05 < ocoll["accepts"].mean() < 0.95
    else:
        nrm, uni = rng.normal(size=(T, C, d + 1)), rng.random(size=(T, C, 52))
        coll, last = s.run(st, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)))
        ost = oa.asss_init(pot, q0)
        ost = ost._replace(adapt_state=ost.adapt_state._replace(scale=np.broadcast_to(np.eye(d) * 0.02, (C, d, d)).copy()))
        olast, _ = oa.asss_run(ost, pot, T, draws=(nrm, uni), num_warmup=5)
    np.testing.assert_allclose(last.z["beta"].cpu().numpy(), olast.z, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(last.adapt_state.scale.cpu().numpy(), olast.adapt_state.scale, rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(last.potential_energy.cpu().numpy(), olast.potential_energy, rtol=1e-9)
