"""Pins the CPU oracle (oracle/) -- the checker every GPU parity test relies on -- against:
the reference's recorded notebook outputs (tests/golden/reference_pins.json), independent SciPy
formulas, the algebraic contract of cholesky_update, Random123 known answers, and itself
(NumPy restatement vs C restatement)."""
import json
import math
import os

import numpy as np
import pytest
import scipy.optimize
import scipy.stats as ss

from oracle import arwmh_numpy as o
from oracle import c_oracle as co

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_pins.json")))


# ---- potentials ---------------------------------------------------------------------------
def _es_scipy(q):
    y, sg = o.EIGHT_SCHOOLS_Y, o.EIGHT_SCHOOLS_SIGMA
    mu, t, eta = q[0], q[1], q[2:]
    tau = math.exp(t)
    lp = ss.norm.logpdf(mu, 0, 5) + ss.halfcauchy.logpdf(tau, scale=5) + t
    lp += ss.norm.logpdf(eta).sum() + ss.norm.logpdf(y, mu + tau * eta, sg).sum()
    return -lp


def test_eight_schools_known_answers():
    # SURVEY 8c known answers (hand formula of the reference model, checked against the energy pins)
    q = np.zeros((1, 10))
    assert o.potential_eight_schools(q)[0] == pytest.approx(43.43563727714813, abs=1e-10)
    q = np.array([[4.4, math.log(2.8), .32, .08, -.09, .05, -.16, -.09, .36, .08]])
    assert o.potential_eight_schools(q)[0] == pytest.approx(41.43286387717417, abs=1e-10)


def test_eight_schools_vs_scipy_and_energy_pin():
    rng = np.random.default_rng(0)
    q = rng.normal(size=(20, 10))
    ref = np.array([_es_scipy(r) for r in q])
    np.testing.assert_allclose(o.potential_eight_schools(q), ref, rtol=1e-12)
    np.testing.assert_allclose(co.potential("eight_schools", q), ref, rtol=1e-12)
    # the unconstrained minimum must lie below every potential energy the reference ever recorded
    res = scipy.optimize.minimize(lambda v: o.potential_eight_schools(v[None])[0], np.zeros(10), method="BFGS")
    pin = GOLD["energy_pins"]["eight_schools_min_U_100x1e6_steps"][0]
    assert res.fun < pin < res.fun + 1.0
    assert GOLD["eight_schools_arwmh_table"]["min_potential_energy"] > pin


def _diamonds_scipy(q, X, Y):
    Xc = X[:, 1:] - X[:, 1:].mean(0)
    icpt, b, s = q[0], q[1:-1], q[-1]
    sig = math.exp(s)
    lp = ss.norm.logpdf(b).sum() + ss.t.logpdf(icpt, 3, 8, 10)
    lp += math.log(2) + ss.t.logpdf(sig, 3, 0, 10) + s
    lp += ss.norm.logpdf(Y, icpt + Xc @ b, sig).sum()
    return -lp


def test_diamonds_vs_scipy():
    from adaptive_mcmc_b200.models import synthetic_diamonds
    data = synthetic_diamonds(n=300, k=25, seed=1)
    rng = np.random.default_rng(1)
    q = rng.normal(size=(6, 26)) * 0.3
    q[:, 0] += 7.8
    q[:, -1] = rng.normal(size=6) * 0.2 - 2.0
    ref = np.array([_diamonds_scipy(r, data["X"], data["Y"]) for r in q])
    np.testing.assert_allclose(o.potential_diamonds(q, data["X"], data["Y"]), ref, rtol=1e-11)
    np.testing.assert_allclose(co.potential("diamonds", q, **data), ref, rtol=1e-11)


def test_kidiq_vs_scipy():
    from adaptive_mcmc_b200.models import synthetic_kidiq
    data = synthetic_kidiq()
    rng = np.random.default_rng(2)
    q = np.column_stack([rng.normal(26, 2, 5), rng.normal(6, 1, 5), rng.normal(0.55, 0.05, 5), rng.normal(2.9, 0.1, 5)])
    ref = []
    for r in q:
        sig = math.exp(r[3])
        mu = r[0] + r[1] * data["mom_hs"] + r[2] * data["mom_iq"]
        ref.append(-(ss.halfcauchy.logpdf(sig, scale=2.5) + r[3] + ss.norm.logpdf(data["kid_score"], mu, sig).sum()))
    np.testing.assert_allclose(o.potential_kidiq(q, **data), ref, rtol=1e-12)
    np.testing.assert_allclose(co.potential("kidiq", q, **data), ref, rtol=1e-12)


def test_gaussian_potential():
    from adaptive_mcmc_b200.models import ar1_precision_chol
    P = ar1_precision_chol(12, 0.9)
    rng = np.random.default_rng(3)
    q = rng.normal(size=(4, 12))
    Q = P @ P.T
    ref = 0.5 * np.einsum("ci,ij,cj->c", q, Q, q)
    np.testing.assert_allclose(o.potential_gaussian(q, P), ref, rtol=1e-12)
    np.testing.assert_allclose(co.potential("gaussian", q, prec_chol=P), ref, rtol=1e-12)


# ---- cholesky_update ------------------------------------------------------------------------
@pytest.mark.parametrize("d", [1, 2, 10, 26])
@pytest.mark.parametrize("coef", [0.3, 1.0, -0.05])
def test_cholesky_update_contract(d, coef):
    rng = np.random.default_rng(d)
    A = rng.normal(size=(5, d, d))
    L = np.linalg.cholesky(A @ A.transpose(0, 2, 1) + np.eye(d))
    x = rng.normal(size=(5, d))
    if coef < 0:
        x = x * 0.5  # keep the downdate positive definite
    Ln = o.cholesky_update(L, x, coef)
    target = L @ L.transpose(0, 2, 1) + coef * x[:, :, None] * x[:, None, :]
    assert np.abs(Ln @ Ln.transpose(0, 2, 1) - target).max() / np.abs(target).max() < 1e-12
    assert np.abs(np.triu(Ln, 1)).max() == 0.0
    assert (np.einsum("cii->ci", Ln) > 0).all()
    np.testing.assert_allclose(Ln, np.linalg.cholesky(target), rtol=1e-9, atol=1e-11)


def test_cholesky_update_zero_diagonal_is_nan():
    # gamma == 1 => sqrt(1-gamma) L == 0 => NaN => the reference keeps the old factor (arwmh.py:191)
    L = np.zeros((2, 3, 3))
    assert np.isnan(o.cholesky_update(L, np.ones((2, 3)), 1.0)).any(axis=(1, 2)).all()
    pot = o.make_potential("eight_schools")
    st = o.arwmh_init(pot, np.zeros((3, 10)))
    rng = np.random.default_rng(0)
    new, _, _ = o.arwmh_step(st, pot, rng.normal(size=(3, 10)), rng.random(3))
    np.testing.assert_array_equal(new.adapt_state.scale, st.adapt_state.scale)  # kept at n == 1
    new2, _, _ = o.arwmh_step(new, pot, rng.normal(size=(3, 10)), rng.random(3))
    assert not np.array_equal(new2.adapt_state.scale, new.adapt_state.scale)


# ---- RNG ------------------------------------------------------------------------------------
def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, exp in kat:
        got = tuple(int(v) for v in o.philox4x32_10(*ctr, *key))
        assert got == exp


def test_philox_draw_moments():
    z, u = o.philox_draws(1, np.arange(200000), 5, 10)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert abs(u.mean() - 0.5) < 3e-3 and u.min() >= 0 and u.max() < 1
    assert abs(ss.kurtosis(z.ravel())) < 0.02


# ---- NumPy restatement vs C restatement ------------------------------------------------------
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-10), (np.float32, 2e-3)])
def test_numpy_vs_c_oracle(dt, tol):
    rng = np.random.default_rng(0)
    C, T = 12, 150
    q0 = co.init_uniform(3, C, 10, dt=dt)
    np.testing.assert_array_equal(q0, o.philox_init_uniform(3, np.arange(C), 10, dt=dt))
    pot = o.make_potential("eight_schools")
    st = o.arwmh_init(pot, q0)
    nrm = rng.normal(size=(T, C, 10)).astype(dt)
    uni = rng.random(size=(T, C)).astype(dt)
    for kw in (dict(), dict(num_warmup=40, lr_decay=0.5), dict(adapt=False)):
        s1, o1 = o.arwmh_run(st, pot, T, draws=(nrm, uni), record_accept=True, thinning=3, collect_start=2, **kw)
        s2, o2 = co.arwmh_run(st, "eight_schools", T, draws=(nrm, uni), record_accept=True, thinning=3, collect_start=2, **kw)
        same = (o1["accepts"] == o2["accepts"]).all(axis=0)
        assert same.mean() >= (1.0 if dt == np.float64 else 0.8)
        assert o1["z"].shape == o2["z"].shape == ((T - 2) // 3, C, 10)
        np.testing.assert_allclose(o1["z"][:, same], o2["z"][:, same], rtol=tol, atol=tol)
        np.testing.assert_allclose(s1.adapt_state.scale[same], s2.adapt_state.scale[same], rtol=tol, atol=tol)
        np.testing.assert_allclose(s1.as_change[same], s2.as_change[same], rtol=tol, atol=tol)


def test_c_oracle_other_models():
    from adaptive_mcmc_b200.models import synthetic_kidiq, synthetic_diamonds
    rng = np.random.default_rng(5)
    for model, data, d in (("kidiq", synthetic_kidiq(), 4), ("diamonds", synthetic_diamonds(n=200), 26),
                           ("std_normal", dict(), 2)):
        pot = o.make_potential(model, **data)
        q0 = rng.uniform(-2, 2, size=(4, d))
        st = o.arwmh_init(pot, q0)
        T = 60
        nrm = rng.normal(size=(T, 4, d)); uni = rng.random(size=(T, 4))
        s1, o1 = o.arwmh_run(st, pot, T, draws=(nrm, uni), record_accept=True)
        s2, o2 = co.arwmh_run(st, model, T, draws=(nrm, uni), record_accept=True, **data)
        assert (o1["accepts"] == o2["accepts"]).all()
        np.testing.assert_allclose(s1.z, s2.z, rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(s1.adapt_state.scale, s2.adapt_state.scale, rtol=1e-8, atol=1e-10)


# ---- statistical pin: the oracle sampler reproduces the reference's recorded posterior table ----
def test_oracle_reproduces_reference_posterior_table():
    tab = GOLD["eight_schools_arwmh_table"]
    C, warm, T = 16, 20000, 120000
    q0 = co.init_uniform(11, C, 10, dt=np.float32)
    st = o.arwmh_init(o.make_potential("eight_schools"), q0)
    last, out = co.arwmh_run(st, "eight_schools", warm + T, seed=11, num_warmup=warm, thinning=20, collect_start=warm)
    z = out["z"].astype(np.float64)  # [S, C, 10]
    z[..., 1] = np.exp(z[..., 1])    # tau
    mean = z.mean(axis=(0, 1)); std = z.std(axis=(0, 1))
    # Monte-Carlo error of these averages is ~0.03; tau is heavy-tailed (looser)
    np.testing.assert_allclose(mean[[0] + list(range(2, 10))], np.array(tab["mean"])[[0] + list(range(2, 10))], atol=0.12)
    np.testing.assert_allclose(std[[0] + list(range(2, 10))], np.array(tab["std"])[[0] + list(range(2, 10))], atol=0.12)
    assert abs(mean[1] - tab["mean"][1]) < 0.35 and abs(std[1] - tab["std"][1]) < 0.6
    assert out["potential_energy"].min() > 40.05
    acc = last.mean_accept_prob.mean()
    assert 0.18 < acc < 0.30  # Robbins-Monro drives acceptance to 0.234


# ---- drivers / diagnostics -------------------------------------------------------------------
def test_ns_logscale_grid():
    g = o.ns_logscale(6)
    assert len(g) == 460 and g[0] == 1 and g[-1] == 10**6
    assert list(g[:12]) == list(range(1, 13))
    assert g[100] == 110 and g[190] == 1100


def test_ess_and_rhat():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(4, 5000))
    ess = o.effective_sample_size(x)
    assert 0.85 * 20000 < ess < 1.15 * 20000
    rho = 0.8
    y = np.zeros((4, 20000))
    e = rng.normal(size=y.shape)
    for t in range(1, y.shape[1]):
        y[:, t] = rho * y[:, t - 1] + e[:, t]
    ess = o.effective_sample_size(y)
    expect = 80000 * (1 - rho) / (1 + rho)
    assert 0.8 * expect < ess < 1.2 * expect
    assert abs(o.split_gelman_rubin(x) - 1) < 0.01
    assert o.split_gelman_rubin(x + np.arange(4)[:, None]) > 1.3


# ---- sample-quality metrics oracle (python/utils/evaluation.py) --------------------------------------------
def test_evaluation_oracle_identities():
    from oracle import evaluation_numpy as oe
    from scipy.stats import wasserstein_distance

    rng = np.random.default_rng(0)
    x = rng.normal(size=(200, 5)).astype(np.float32)
    y = (rng.normal(size=(150, 5)) + 0.5).astype(np.float32)
    # independent formula for the kernel sum: |x|^2 + |y|^2 - 2 x.y in float64
    d2 = (x.astype(np.float64) ** 2).sum(1)[:, None] + (y.astype(np.float64) ** 2).sum(1)[None] - 2 * x.astype(np.float64) @ y.astype(np.float64).T
    np.testing.assert_allclose(oe.sqdist(x, y), d2, rtol=2e-5, atol=1e-5)
    assert oe.mmd_heuristic(x, x) < 1e-3
    assert oe.mmd_heuristic(x, y) > 0.1
    # mmd2_unbiased has zero mean under the null: two halves of one sample give a value near 0 (either sign)
    assert abs(oe.mmd2_unbiased(x[:100], x[100:], 0.5)) < 0.02
    # 1-D Wasserstein against SciPy
    a, b = x[:, 0], y[:200 - 50, 0]
    assert abs(float(oe.wasserstein_1d(a[:150], b, 1.0)) - wasserstein_distance(a[:150], b)) < 1e-6
    # the 1-1 coupling of a sample with a shifted copy of itself costs exactly the shift
    assert abs(oe.wasserstein_dist11_p(x[:60], x[:60] + np.float32(0.25), 2.0) - 0.25 * np.sqrt(5)) < 1e-5
    assert abs(oe.pth_moment_rmse(x, x, 2.0)) == 0.0


def test_auction_restatement_matches_scipy():
    """oracle/assignment_numpy.py (the algorithm of csrc/assign.cu) ends at SciPy's optimum on the integer-scaled costs."""
    from scipy.optimize import linear_sum_assignment

    from oracle import assignment_numpy as oa

    rng = np.random.default_rng(4)
    for n, d in ((1, 2), (2, 2), (9, 3), (80, 6), (250, 26)):
        x, y = rng.normal(size=(n, d)), rng.normal(size=(n, d)) + 0.2
        ci = oa.quantise(np.linalg.norm(x[:, None] - y[None], axis=-1).astype(np.float32))
        col, _ = oa.auction(ci)
        ri, cj = linear_sum_assignment(ci)
        assert sorted(col.tolist()) == list(range(n)) and ci[np.arange(n), col].sum() == ci[ri, cj].sum()
    ci = oa.quantise(rng.integers(0, 3, size=(40, 40)).astype(np.float32))
    col, _ = oa.auction(ci)
    ri, cj = linear_sum_assignment(ci)
    assert ci[np.arange(40), col].sum() == ci[ri, cj].sum()
