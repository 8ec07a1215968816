"""GPU tests of the tcgen05 tensor-core path (building blocks first, then the diamonds kernel)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [16, 80, 96])
@pytest.mark.parametrize("use_tma", [0, 1])
def test_umma_selftest(K, use_tma):
    g = torch.Generator().manual_seed(K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(256, K, generator=g).to(torch.bfloat16).cuda()
    D = torch.zeros(128, 256, dtype=torch.float32, device="cuda")
    scratch = torch.zeros(256 * K, dtype=torch.int16, device="cuda")
    swap = int(os.environ.get("AMCMC_UMMA_SWAP", "0"))
    rc = _lib.lib().amcmc_selftest_umma(A.data_ptr(), B.data_ptr(), K, scratch.data_ptr(), D.data_ptr(), use_tma, swap,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "amcmc_selftest_umma")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (D - ref).abs().max().item()
    assert err < 1e-3 * (1 + ref.abs().max().item()), f"max err {err}"


# ---- diamonds on the tensor cores: frozen / pooled kernel vs the float64 oracle ----------------------
from adaptive_mcmc_b200 import models
from adaptive_mcmc_b200.parallel import PooledARWMH
from oracle import arwmh_numpy as o
from oracle import pooled_numpy as op


@pytest.fixture(scope="module")
def diamonds_data():
    return models.synthetic_diamonds(n=5000, k=25, seed=0)


def _mode(data):
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    return np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])


@pytest.mark.parametrize("C", [128, 1000])
def test_diamonds_tc_logdensity_and_trajectory(C, diamonds_data):
    """Shared draws: the tcgen05 path must take the same accept decisions as the fp64 oracle, and every
    potential energy it stores must equal the fp64 potential of the stored position (abs 1e-3 on
    U ~ -3.3e3, i.e. 3e-7 relative -- the fp32 reference's own resolution there is 2.4e-4)."""
    d, T = 26, 12
    rng = np.random.default_rng(1)
    q0 = _mode(diamonds_data)[None] + 0.004 * rng.normal(size=(C, d))
    s = PooledARWMH(models.diamonds, num_chains=C, pool_every=T, init_strategy=am.init_to_value(torch.from_numpy(q0)),
                    impl=_lib.IMPL_TENSOR)
    s.init(7, model_kwargs=diamonds_data)
    # a sensible shared proposal: scale ~ posterior sd
    s.scale.mul_(0.002)
    pot = o.make_potential("diamonds", **diamonds_data)
    z0 = s.batch.z.t().double().cpu().numpy()
    U0 = pot(z0)
    np.testing.assert_allclose(s.batch.pe.cpu().numpy(), U0, rtol=2e-5)
    pool = op.pooled_init(z0)
    pool["loc"] = s.loc.double().cpu().numpy()
    pool["L"] = s.dense_scale().double().cpu().numpy()
    nrm = rng.normal(size=(T, C, d)).astype(np.float32)
    uni = rng.random(size=(T, C)).astype(np.float32)
    raw = s.run_window(T, draws=(torch.from_numpy(nrm).permute(0, 2, 1).contiguous(), torch.from_numpy(uni)),
                       record_accept=True, adapt=True)
    zo, Uo, pool2, info = op.pooled_window(pot, z0, U0, pool, T, 0, draws=(nrm.astype(np.float64), uni.astype(np.float64)),
                                           record=True)
    acc_g = raw["accept"].cpu().numpy().astype(bool)
    same = (acc_g == info["accepts"]).all(axis=0)
    assert same.mean() > 0.97, same.mean()
    assert 0.05 < acc_g.mean() < 0.8
    zg = raw["z"].permute(0, 2, 1).cpu().numpy().astype(np.float64)  # [T, C, d]
    peg = raw["potential_energy"].cpu().numpy().astype(np.float64)
    # (1) stored energy == fp64 potential of the stored position
    Uchk = pot(zg.reshape(-1, d)).reshape(T, C)
    moved = np.cumsum(acc_g, axis=0) > 0          # energy produced by the tensor-core path
    err = np.abs(peg - Uchk)
    assert err[moved].max() < 1e-3, err[moved].max()
    # energies still equal to U0 come from the fp32 CUDA-core init kernel, whose residual Y - I - Xb cancels
    # ~6 digits in fp32 exactly as the fp32 reference does (the centred tensor-core path does not)
    assert err[~moved].max() < 1e-2, err[~moved].max()
    # (2) trajectories of the chains with identical decisions
    assert np.abs(zg[:, same] - info["z"][:, same]).max() < 2e-5
    # (3) pooled update: mean acceptance, loc, cov -> scale, log_step_size
    np.testing.assert_allclose(s.batch.macc.cpu().numpy()[same], info["mean_accept"][same], atol=2e-3)
    if same.all():
        np.testing.assert_allclose(s.loc.cpu().numpy(), pool2["loc"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(s.cov.cpu().numpy(), pool2["cov"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(s.dense_scale().cpu().numpy(), pool2["L"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(float(s.log_step_size), pool2["lam"], atol=1e-4)
    assert s.window == 1


def test_diamonds_tc_far_from_reference_point(diamonds_data):
    """Chains far from q_ref (the transient): Delta is large, the split-bf16 GEMM must still give
    fp32-class RELATIVE accuracy on the potential."""
    C, d, T = 256, 26, 4
    rng = np.random.default_rng(2)
    q0 = rng.uniform(-2, 2, size=(C, d))
    s = PooledARWMH(models.diamonds, num_chains=C, pool_every=T, init_strategy=am.init_to_value(torch.from_numpy(q0)),
                    impl=_lib.IMPL_TENSOR)
    s.init(3, model_kwargs=diamonds_data)
    s.scale.mul_(0.05)
    raw = s.run_window(T, adapt=False)
    pot = o.make_potential("diamonds", **diamonds_data)
    zg = raw["z"].permute(0, 2, 1).cpu().numpy().astype(np.float64)
    peg = raw["potential_energy"].cpu().numpy().astype(np.float64)
    Uchk = pot(zg.reshape(-1, d)).reshape(T, C)
    assert (np.abs(peg - Uchk) / np.abs(Uchk)).max() < 2e-5


def test_diamonds_tc_matches_cuda_core_frozen_path(diamonds_data):
    """Same frozen run through the tensor-core kernel and through the exact CUDA-core block kernel
    (Philox draws): identical accept decisions for almost every chain."""
    C, d, T = 512, 26, 10
    rng = np.random.default_rng(4)
    q0 = _mode(diamonds_data)[None] + 0.004 * rng.normal(size=(C, d))
    out = {}
    for impl in (_lib.IMPL_TENSOR, _lib.IMPL_BLOCK):
        s = PooledARWMH(models.diamonds, num_chains=C, pool_every=T, init_strategy=am.init_to_value(torch.from_numpy(q0)), impl=impl)
        s.init(5, model_kwargs=diamonds_data)
        s.scale.mul_(0.002)
        raw = s.run_window(T, record_accept=True, adapt=False)
        out[impl] = (raw["accept"].cpu().numpy(), s.batch.z.clone(), s.batch.pe.clone())
    same = (out[_lib.IMPL_TENSOR][0] == out[_lib.IMPL_BLOCK][0]).all(axis=0)
    assert same.mean() > 0.97
    dz = (out[_lib.IMPL_TENSOR][1] - out[_lib.IMPL_BLOCK][1]).abs().cpu().numpy()[:, same]
    assert dz.max() < 2e-5


def test_pooled_adaptation_converges_on_diamonds(diamonds_data):
    """BASELINE.json configs[3] in miniature: 4096 chains, pooled adaptation every 100 steps; the shared
    covariance converges to the analytic posterior covariance of the regression coefficients."""
    C, d = 4096, 26
    rng = np.random.default_rng(6)
    q0 = _mode(diamonds_data)[None] + 0.01 * rng.normal(size=(C, d))
    s = PooledARWMH(models.diamonds, num_chains=C, pool_every=100, init_strategy=am.init_to_value(torch.from_numpy(q0)))
    s.init(9, model_kwargs=diamonds_data)
    s.scale.mul_(0.01)
    s.cov.mul_(1e-4)
    coll = s.run(6000, thinning=100)
    X, Y = diamonds_data["X"], diamonds_data["Y"]
    Xc = X[:, 1:] - X[:, 1:].mean(0)
    sig = float(torch.exp(coll["z"]["sigma"][-10:]).mean())
    assert abs(sig - 0.123) < 0.004
    post_sd = np.sqrt(np.diag(sig**2 * np.linalg.inv(Xc.T @ Xc + sig**2 * np.eye(24))))
    est_sd = np.sqrt(np.diag(s.cov.cpu().numpy()))[1:25]
    assert 0.7 < (est_sd / post_sd).min() and (est_sd / post_sd).max() < 1.4, est_sd / post_sd
    acc = float(s.batch.macc.mean())
    assert 0.15 < acc < 0.35, acc
    b_last = coll["z"]["b"][-20:].double().mean((0, 1)).cpu().numpy()
    ridge = np.linalg.solve(Xc.T @ Xc + sig**2 * np.eye(24), Xc.T @ (Y - Y.mean()))
    assert (np.abs(b_last - ridge) / post_sd).max() < 0.5


# ---- diamonds on the tensor cores WITH per-chain adaptation (the full ARWMH.sample) ------------------------
from oracle import c_oracle as co


def _adaptive_sampler(C, q0, impl, **kw):
    s = am.ARWMH(models.diamonds, num_chains=C, init_strategy=am.init_to_value(torch.from_numpy(q0)), **kw)
    s.impl = impl
    return s


@pytest.mark.parametrize("C", [256, 1000])
def test_diamonds_tc_adaptive_matches_oracle(C, diamonds_data):
    """Shared draws, per-chain adaptation: tensor-core kernel (fp32 state, split-bf16 likelihood) vs the fp64 oracle
    of the reference step.  Same accept decisions for almost every chain; positions, running mean, factor and step
    size of those chains agree."""
    d, T, nw = 26, 40, 10
    rng = np.random.default_rng(3)
    q0 = _mode(diamonds_data)[None] + 0.004 * rng.normal(size=(C, d))
    s = _adaptive_sampler(C, q0, _lib.IMPL_TENSOR)
    st = s.init(1, num_warmup=nw, init_params=None, model_kwargs=diamonds_data)
    # a sensible start for the factor: the reference starts from I, which rejects everything on this posterior
    b = am.ChainBatch.from_state(s.potential, st)
    b.set_dense_scale(torch.eye(d) * 0.002)
    st = b.to_state()
    pot = o.make_potential("diamonds", **diamonds_data)
    z0 = b.z.t().double().cpu().numpy()
    ost = o.ARWMHState(0, z0, pot(z0), np.zeros(C), o.ARWMHAdaptState(z0.copy(), np.broadcast_to(np.eye(d) * 0.002, (C, d, d)).copy(),
                                                                     np.zeros(C)), np.zeros(C), 0)
    nrm = rng.normal(size=(T, C, d)).astype(np.float32)
    uni = rng.random(size=(T, C)).astype(np.float32)
    coll, last = s.run(st, T, draws=(torch.from_numpy(nrm), torch.from_numpy(uni)), record_accept=True)
    olast, ocoll = co.arwmh_run(ost, "diamonds", T, draws=(nrm.astype(np.float64), uni.astype(np.float64)), record_accept=True,
                                num_warmup=nw, **diamonds_data)
    acc_g = coll["accept"].cpu().numpy()
    same = (acc_g == ocoll["accepts"]).all(axis=0)
    assert same.mean() > 0.93, same.mean()
    assert 0.05 < acc_g.mean() < 0.9
    # the chains that differ are the ones whose u fell inside the float32 energy error of alpha -- and only those
    import flipcheck
    orc, _ = flipcheck.oracle_steps(o, ost, pot, nrm, uni, num_warmup=nw)
    flipcheck.analyse(acc_g, coll["potential_energy"].cpu().numpy(), orc, uni, f"diamonds tcgen05 adaptive C={C}")
    zg = np.concatenate([v.cpu().numpy().reshape(T, C, -1) for v in coll["z"].values()], axis=-1).astype(np.float64)
    # fp32 state vs the fp64 oracle: the north star's fp32 bar (1e-3 relative), in practice ~1e-5
    assert (np.abs(zg[:, same] - ocoll["z"][:, same]) / (1 + np.abs(ocoll["z"][:, same]))).max() < 1e-3
    a, oa_ = last.adapt_state, olast.adapt_state
    # The posterior is sharp (sd ~ 0.002 per coordinate) and the step is chaotic in the energies: the ~1e-3 absolute
    # error of fp32 energies enters lambda through gamma*(alpha - target), lambda scales the next proposal, and the
    # difference grows ~40x over these 40 steps (scripts/probes/tc_adapt_diag.py prints the growth curve).  So the adapted
    # state is compared chain by chain: nearly all chains inside the tight band, every chain inside a loose one.
    def close(x, y, rtol, atol):
        x = np.asarray(x.cpu().numpy() if hasattr(x, "cpu") else x, np.float64)[same].reshape(int(same.sum()), -1)
        y = np.asarray(y, np.float64)[same].reshape(int(same.sum()), -1)
        return (np.abs(x - y) <= atol + rtol * np.abs(y)).all(axis=1)

    def fro(x, y):  # relative Frobenius distance of the factors (elementwise is meaningless for the near-zero entries)
        x = x.cpu().numpy().astype(np.float64)[same]
        return np.sqrt(((x - y[same]) ** 2).sum((1, 2)) / (y[same] ** 2).sum((1, 2)))

    checks = {
        "loc": close(a.loc, oa_.loc, 1e-3, 1e-4),
        "log_step_size": close(a.log_step_size, oa_.log_step_size, 0, 3e-3),
        "scale": fro(a.scale, oa_.scale) < 1e-2,
        "mean_accept_prob": close(last.mean_accept_prob, olast.mean_accept_prob, 0, 3e-3),
        "as_change": close(last.as_change, olast.as_change, 5e-2, 1e-6),
        "potential_energy": close(last.potential_energy, olast.potential_energy, 0, 3e-2),
    }
    frac = {k: float(v.mean()) for k, v in checks.items()}
    assert all(f > 0.95 for f in frac.values()), frac
    assert close(a.log_step_size, oa_.log_step_size, 0, 0.1).all()
    assert close(a.loc, oa_.loc, 1e-2, 1e-3).all()
    assert (fro(a.scale, oa_.scale) < 0.2).all()
    assert int(last.i) == T


@pytest.mark.parametrize("C,max_ctas", [(512, 0), (512, 1), (1000, 1), (1000, 3), (700, 2)])
def test_diamonds_tc_adaptive_matches_block_kernel(C, max_ctas, diamonds_data, monkeypatch):
    """Philox draws: the tensor-core adaptive path and the exact CUDA-core block kernel run the same chains.
    `max_ctas` caps the grid (AMCMC_TC_MAX_CTAS) so that a CTA serves 2-4 chain groups, in one or two rounds, in
    two alternating streams -- the situation of 65,536 chains on 148 SMs -- at a chain count the test can afford."""
    if max_ctas:
        monkeypatch.setenv("AMCMC_TC_MAX_CTAS", str(max_ctas))
    d, T = 26, 30
    rng = np.random.default_rng(8)
    q0 = _mode(diamonds_data)[None] + 0.004 * rng.normal(size=(C, d))
    res = {}
    for impl in (_lib.IMPL_TENSOR, _lib.IMPL_BLOCK):
        s = _adaptive_sampler(C, q0, impl)
        st = s.init(2, num_warmup=0, init_params=None, model_kwargs=diamonds_data)
        b = am.ChainBatch.from_state(s.potential, st)
        b.set_dense_scale(torch.eye(d) * 0.002)
        raw = s.run_batch(b, T, thinning=5, record_accept=True)
        # cut the same stream into two launches (state round trip through the ABI layout)
        raw2 = s.run_batch(b, 10, collect=())
        res[impl] = (raw["accept"].cpu().numpy(), b.z.clone(), b.scale.clone(), b.lam.clone(), raw["z"].clone())
    same = (res[_lib.IMPL_TENSOR][0] == res[_lib.IMPL_BLOCK][0]).all(axis=0)
    assert same.mean() > 0.9, same.mean()
    sel = torch.from_numpy(same).to(res[_lib.IMPL_TENSOR][1].device)
    ref = res[_lib.IMPL_BLOCK][4][:, :, sel]
    assert ((res[_lib.IMPL_TENSOR][4][:, :, sel] - ref).abs() / (1 + ref.abs())).max() < 1e-3


def test_diamonds_tc_adaptive_converges(diamonds_data):
    """Long run of the per-chain-adaptive tensor-core path: 8192 chains x 20,000 steps, every chain adapting on its
    own.  The cross-chain sample reproduces the analytic posterior of the regression coefficients, acceptance sits
    at the Robbins-Monro target, and each chain's own adapted mean is inside the posterior."""
    C, d = 8192, 26
    rng = np.random.default_rng(12)
    q0 = _mode(diamonds_data)[None] + 0.004 * rng.normal(size=(C, d))
    s = _adaptive_sampler(C, q0, _lib.IMPL_AUTO)  # auto dispatch must pick the tensor-core path at this chain count
    st = s.init(3, num_warmup=0, init_params=None, model_kwargs=diamonds_data)
    b = am.ChainBatch.from_state(s.potential, st)
    b.set_dense_scale(torch.eye(d) * 0.002)
    s.run_batch(b, 15000, collect=())
    raw = s.run_batch(b, 5000, thinning=500, collect=("z", "potential_energy"))
    assert int(b.i) == 20000
    z = raw["z"].double()  # [S, d, C]
    X, Y = diamonds_data["X"], diamonds_data["Y"]
    Xc = X[:, 1:] - X[:, 1:].mean(0)
    sig = float(torch.exp(z[:, 25, :]).mean())
    assert abs(sig - 0.123) < 0.003, sig
    post_cov = sig**2 * np.linalg.inv(Xc.T @ Xc + sig**2 * np.eye(24))
    post_sd = np.sqrt(np.diag(post_cov))
    ridge = np.linalg.solve(Xc.T @ Xc + sig**2 * np.eye(24), Xc.T @ (Y - Y.mean()))
    bsamp = z[:, 1:25, :].permute(0, 2, 1).reshape(-1, 24).cpu().numpy()
    # 10 thinned states x 8192 chains: mean within a few MCSE, sd ratio near 1.  Columns 1-4 of the design are
    # collinear (rho up to 0.999): adaptation of those directions is the slow part of this posterior
    assert (np.abs(bsamp.mean(0) - ridge) / post_sd).max() < 0.25, np.abs(bsamp.mean(0) - ridge) / post_sd
    ratio = bsamp.std(0) / post_sd
    # the collinear columns have posterior sds far above the 0.002 start of the factor: 20k steps of Robbins-Monro
    # adaptation have not opened those directions yet (a property of the algorithm: the exact block kernel shows the
    # same spreads, see test_diamonds_tc_adaptive_spread_matches_block_kernel), so they are only required not to overshoot
    assert 0.8 < ratio[4:].min() and ratio.max() < 1.25, ratio
    acc = b.macc.cpu().numpy()
    assert abs(acc.mean() - 0.234) < 0.03, acc.mean()
    # every chain's adapted mean lies inside the posterior (|loc - ridge| below ~6 sd in every coordinate)
    loc = b.loc.t().double().cpu().numpy()[:, 1:25]
    assert (np.abs(loc - ridge) / post_sd).max() < 8.0
    assert torch.isfinite(b.scale).all() and torch.isfinite(b.lam).all()


def test_diamonds_tc_adaptive_spread_matches_block_kernel(diamonds_data):
    """Independent long runs (own Philox streams) of the tensor-core adaptive path and of the exact CUDA-core block
    kernel: the cross-chain spread of every coefficient after 20,000 steps agrees, including the slowly adapting
    collinear directions."""
    C, d = 1024, 26
    rng = np.random.default_rng(13)
    q0 = _mode(diamonds_data)[None] + 0.004 * rng.normal(size=(C, d))
    sd = {}
    for impl in (_lib.IMPL_TENSOR, _lib.IMPL_BLOCK):
        s = _adaptive_sampler(C, q0, impl)
        st = s.init(4, num_warmup=0, init_params=None, model_kwargs=diamonds_data)
        b = am.ChainBatch.from_state(s.potential, st)
        b.set_dense_scale(torch.eye(d) * 0.002)
        s.run_batch(b, 18000, collect=())
        raw = s.run_batch(b, 2000, thinning=500, collect=("z",))
        sd[impl] = raw["z"].double().permute(0, 2, 1).reshape(-1, d).std(0).cpu().numpy()
    r = sd[_lib.IMPL_TENSOR] / sd[_lib.IMPL_BLOCK]
    assert 0.85 < r.min() and r.max() < 1.18, r


@pytest.mark.parametrize("segment", [1, 7, 16])
def test_diamonds_tc_adaptive_segments(segment, diamonds_data, monkeypatch):
    """Long launches are cut into segments with a re-centred GEMM reference point (256 steps by default).  Forcing
    short segments must give the same chains (same Philox streams, same accept decisions up to the energy round-off)
    and exactly the same sample bookkeeping: thinning, collect_start, accept record, iteration counter."""
    C, d, T, thin, skip = 384, 26, 41, 3, 4
    rng = np.random.default_rng(21)
    q0 = _mode(diamonds_data)[None] + 0.004 * rng.normal(size=(C, d))
    res = []
    for seg in (None, segment):
        if seg is None:
            monkeypatch.delenv("AMCMC_TC_SEGMENT", raising=False)
        else:
            monkeypatch.setenv("AMCMC_TC_SEGMENT", str(seg))
        s = _adaptive_sampler(C, q0, _lib.IMPL_TENSOR)
        st = s.init(6, num_warmup=10, init_params=None, model_kwargs=diamonds_data)
        b = am.ChainBatch.from_state(s.potential, st)
        b.set_dense_scale(torch.eye(d) * 0.002)
        raw = s.run_batch(b, T, thinning=thin, collect_start=skip, collect=("z", "potential_energy"), record_accept=True)
        res.append((raw, b))
    (r0, b0), (r1, b1) = res
    S = (T - skip) // thin
    assert r0["z"].shape == r1["z"].shape == (S, d, C) and r1["potential_energy"].shape == (S, C)
    assert r1["accept"].shape == (T, C) and int(b1.i) == T
    assert torch.isfinite(r1["z"]).all() and torch.isfinite(r1["potential_energy"]).all()
    same = (r0["accept"] == r1["accept"]).all(dim=0)
    assert same.float().mean() > 0.95
    assert ((r0["z"][:, :, same] - r1["z"][:, :, same]).abs()).max() < 1e-4
    assert ((b0.z[:, same] - b1.z[:, same]).abs()).max() < 1e-4
    assert ((b0.lam[same] - b1.lam[same]).abs()).max() < 5e-2


def test_diamonds_tc_adaptive_energies_from_far_start(diamonds_data):
    """The reference's own start (q0 ~ U(-2,2)^26, factor = I): the batch is spread over |U| ~ 1e4-1e6 for tens of
    thousands of steps.  Every chain has its own GEMM reference point, re-centred every 256 steps, so the stored
    energies keep float32 accuracy there (a reference point shared by the batch measured 0.04 absolute)."""
    C, d = 4096, 26
    s = am.ARWMH(models.diamonds, num_chains=C)
    st = s.init(11, num_warmup=0, init_params=None, model_kwargs=diamonds_data)
    b = am.ChainBatch.from_state(s.potential, st)
    raw = s.run_batch(b, 6000, thinning=500, collect=("z", "potential_energy"))
    pot = o.make_potential("diamonds", **diamonds_data)
    sel = np.arange(0, C, 16)
    z = raw["z"].double().cpu().numpy()[:, :, sel]          # [S, d, 256]
    pe = raw["potential_energy"].double().cpu().numpy()[:, sel]
    exact = np.stack([pot(z[k].T) for k in range(z.shape[0])])
    assert np.abs(exact).max() > 1e4                          # still far from the mode (U ~ -3.3e3 there)
    rel = np.abs(pe - exact) / np.maximum(1.0, np.abs(exact))
    assert rel.max() < 5e-6, rel.max()                        # a few tens of float32 ulps even at |U| ~ 1e6
    mid = np.abs(exact) < 5e4
    err = np.abs(pe - exact)[mid]
    assert mid.any() and err.max() < 0.05 and np.median(err) < 2e-3, (err.max(), np.median(err))
    assert np.median(exact[-1]) < np.median(exact[0])         # and the batch moves downhill
    assert 0.1 < float(b.macc.mean()) < 0.4


def test_diamonds_tc_adaptive_rejects_blown_up_proposals(diamonds_data):
    """NaN / inf potential => reject (arwmh.py:171) on the tensor-core path: draws of 1e30 / inf / nan in a few chains
    poison only those chains' rows of the GEMM; they reject, stay finite, and every other chain matches the exact
    CUDA-core block kernel run with the same draws."""
    C, d, T = 384, 26, 9
    rng = np.random.default_rng(31)
    q0 = _mode(diamonds_data)[None] + 0.004 * rng.normal(size=(C, d))
    nrm = rng.normal(size=(T, C, d)).astype(np.float32)
    uni = rng.random(size=(T, C)).astype(np.float32)
    nrm[2, 0, :] = 1e30
    nrm[3, 130, 5] = np.inf
    nrm[4, 257, 0] = np.nan
    nrm[5, 383, 25] = 1e30      # log sigma -> +inf
    nrm[6, 7, 25] = -1e30       # log sigma -> -inf
    bad = [(2, 0), (3, 130), (4, 257), (5, 383), (6, 7)]
    res = {}
    for impl in (_lib.IMPL_TENSOR, _lib.IMPL_BLOCK):
        s = _adaptive_sampler(C, q0, impl)
        st = s.init(1, num_warmup=0, init_params=None, model_kwargs=diamonds_data)
        b = am.ChainBatch.from_state(s.potential, st)
        b.set_dense_scale(torch.eye(d) * 0.002)
        dr = s._draws_to_device_layout((torch.from_numpy(nrm), torch.from_numpy(uni)))
        raw = s.run_batch(b, T, collect=("z", "potential_energy"), record_accept=True, draws=(dr[0], dr[1].to(dr[0].device)))
        res[impl] = (raw["accept"].cpu().numpy().astype(bool), b.z.clone(), b.pe.clone(), b.scale.clone(), raw["z"].clone())
    acc_t, acc_b = res[_lib.IMPL_TENSOR][0], res[_lib.IMPL_BLOCK][0]
    for (t, c) in bad:
        assert not acc_t[t, c] and not acc_b[t, c]
    for k in range(1, 5):
        assert torch.isfinite(res[_lib.IMPL_TENSOR][k]).all()
    same = (acc_t == acc_b).all(axis=0)
    assert same.mean() > 0.95
    sel = torch.from_numpy(same).to(res[_lib.IMPL_TENSOR][1].device)
    assert (res[_lib.IMPL_TENSOR][1][:, sel] - res[_lib.IMPL_BLOCK][1][:, sel]).abs().max() < 1e-4


def test_diamonds_tc_adaptive_survives_runaway_proposals(diamonds_data):
    """Regression: from the reference's start a chain can jump by hundreds within ten steps (chain 83339 of seed 0:
    log sigma -1.7 -> 0.7 -> 17.7, in the exact block kernel too) and then propose log sigma = -349, where
    e^{-2s} ~ 1e303.  RSS_ref - 2 D.g + sum m^2 cancels catastrophically there; assembled naively its negative value
    times e^{-2s} became U' = -inf, which was accepted and never left.  RSS is now clamped at 0 before the scaling."""
    off, j = 83328, 11
    res = {}
    for impl in (_lib.IMPL_TENSOR, _lib.IMPL_BLOCK):
        s = am.ARWMH(models.diamonds, num_chains=128, chain_offset=off)
        s.impl = impl
        st = s.init(0, num_warmup=0, init_params=None, model_kwargs=diamonds_data)
        b = am.ChainBatch.from_state(s.potential, st, copy=False)
        b.set_dense_scale(torch.eye(26) * 0.002)
        raw = s.run_batch(b, 60, collect=("z", "potential_energy"), record_accept=True)
        res[impl] = (raw["accept"].cpu().numpy().astype(bool), raw["potential_energy"].cpu().numpy(), b)
    acc_t, pe_t, b_t = res[_lib.IMPL_TENSOR]
    acc_b, pe_b, _ = res[_lib.IMPL_BLOCK]
    assert np.isfinite(pe_t).all() and torch.isfinite(b_t.z).all() and torch.isfinite(b_t.scale).all()
    assert np.abs(pe_b[:, j]).max() > 1e5                                                             # this chain does jump
    np.testing.assert_array_equal(acc_t[:12, j], acc_b[:12, j])                                       # incl. the rejection at step 10
    assert (acc_t[:10] == acc_b[:10]).all(axis=0).mean() > 0.9
