"""GPU tests added in round 2 for the paths the round-1 review found untested:

* `amcmc_arwmh_run_host` with ASSS + external draws (normals[T][d+1][C], uniforms[T][52][C]);
* the register kernel's frozen template (`adapt = 0`, ARWMH.sample_Pnx, arwmh.py:230-249) against the oracle on shared draws;
* pooled adaptation on the generic CUDA-core paths (thread-per-chain and CTA-per-chain): the window's mean acceptance
  probability must reach the Robbins-Monro update, so the shared step size converges to the 0.234 target;
* frozen many-chain diamonds `sample_Pnx` runs on the tensor-core shared-state kernel and agrees with the block kernel.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import _lib, models
from adaptive_mcmc_b200.parallel import PooledARWMH
from oracle import arwmh_numpy as o
from oracle import c_oracle as co
from oracle import pooled_numpy as op

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _host_state(batch, host):
    hs = _lib.AmcmcState()
    hs.n_chains, hs.dim, hs.dtype, hs.i = batch.C, batch.d, (0 if batch.z.dtype == torch.float32 else 1), batch.i
    hs.z, hs.potential_energy, hs.mean_accept_prob = host["z"].data_ptr(), host["pe"].data_ptr(), host["macc"].data_ptr()
    hs.loc, hs.scale, hs.log_step_size = host["loc"].data_ptr(), host["scale"].data_ptr(), host["lam"].data_ptr()
    hs.as_change = host["asc"].data_ptr()
    return hs


@pytest.mark.parametrize("T", [7, 1300, 5003])  # several chunks of the host entry (260 samples: 130 + 65 + 33 + 32): the offsets matter
def test_run_host_asss_external_draws(T):
    """The host-buffer entry point must size and offset the external draws by the sampler kind: ASSS reads
    normals[T][d+1][C] and uniforms[T][52][C] (include/amcmc.h).  Bit-identical to the device-pointer path."""
    Cn, d = 96, 10
    s = am.ASSS(models.eight_schools, num_chains=Cn)
    st = s.init(2, num_warmup=0, init_params=None)
    b = s._batch_from_state(st)
    host = {f: getattr(b, f).cpu().pin_memory() for f in b._FIELDS}
    g = torch.Generator().manual_seed(5)
    nrm = torch.randn(T, Cn, d + 1, generator=g)
    uni = torch.rand(T, Cn, 52, generator=g)
    dn, du = s._draws_to_device_layout((nrm, uni))  # [T][d+1][C], [T][52][C]
    # the host entry cuts the run into tapering chunks (amcmc_host_chunk_samples); a launch boundary converts the carried
    # LDL^T factor to the ABI's Cholesky form and back (a rounding), so the device-pointer run is cut at the same steps
    zs, pes, t0, s_left = [], [], 0, T // 5
    while t0 < T:
        ns = _lib.lib().amcmc_host_chunk_samples(s_left, 5) if s_left else 0
        n = ns * 5
        s_left -= ns
        if s_left == 0:  # the last chunk also takes the uncollected tail
            n = T - t0
        raw = s.run_batch(b, n, thinning=5, draws=(dn[t0:t0 + n].contiguous(), du[t0:t0 + n].contiguous()))
        zs.append(raw["z"]); pes.append(raw["potential_energy"])
        t0 += n
    raw = dict(z=torch.cat(zs), potential_energy=torch.cat(pes))
    hn, hu = dn.cpu().contiguous().pin_memory(), du.cpu().contiguous().pin_memory()
    hs = _host_state(b, host)
    hs.i = 0
    a = _lib.AmcmcRunArgs()
    a.n_steps, a.thinning, a.collect_start, a.num_warmup = T, 5, 0, 0
    a.lr_decay, a.target_accept_prob, a.eps = 2 / 3, 0.234, 1e-6
    a.adapt, a.rng_mode, a.seed, a.kernel_kind = 1, _lib.RNG_EXTERNAL, 2, _lib.KERNEL_ASSS
    a.normals, a.uniforms = hn.data_ptr(), hu.data_ptr()
    S = T // 5
    oz = torch.empty(S, d, Cn).pin_memory()
    ope = torch.empty(S, Cn).pin_memory()
    a.out_z, a.out_potential_energy = oz.data_ptr(), ope.data_ptr()
    _lib.check(_lib.lib().amcmc_arwmh_run_host(s.potential.handle, C.byref(hs), C.byref(a)), "run_host")
    assert hs.i == T
    torch.testing.assert_close(oz, raw["z"].cpu(), rtol=0, atol=0)
    torch.testing.assert_close(ope, raw["potential_energy"].cpu(), rtol=0, atol=0)
    torch.testing.assert_close(host["scale"], b.scale.cpu(), rtol=0, atol=0)
    torch.testing.assert_close(host["z"], b.z.cpu(), rtol=0, atol=0)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_frozen_register_kernel_matches_oracle(prec):
    """the `ADAPT = false` template of the thread-per-chain kernel (what ARWMH.sample_Pnx launches) on
    shared draws against the oracle's frozen step (arwmh.py:230-249): identical decisions, positions within tolerance,
    the adaptation state bit-for-bit untouched, mean_accept_prob = mean acceptance probability over the launch."""
    tdt, ndt, tol = (torch.float64, np.float64, 1e-9) if prec == "f64" else (torch.float32, np.float32, 1e-3)
    Cn, d, T = 512, 10, 120
    rng = np.random.default_rng(17)
    s = am.ARWMH(models.eight_schools, num_chains=Cn, dtype=tdt)
    st = s.init(8, num_warmup=0, init_params=None)
    b = am.ChainBatch.from_state(s.potential, st)
    scale0 = np.tril(rng.normal(size=(d, d)) * 0.15 + np.eye(d) * 0.6)
    b.set_dense_scale(torch.from_numpy(scale0))
    b.lam.fill_(-0.4)
    loc0, sc0, lam0 = b.loc.clone(), b.scale.clone(), b.lam.clone()
    nrm = rng.normal(size=(T, Cn, d)).astype(ndt)
    uni = rng.random(size=(T, Cn)).astype(ndt)
    dr = s._draws_to_device_layout((torch.from_numpy(nrm), torch.from_numpy(uni)))
    z0 = _np(b.z.t()).astype(ndt)
    raw = s.run_batch(b, T, draws=dr, adapt=False, record_accept=True)
    assert torch.equal(b.loc, loc0) and torch.equal(b.scale, sc0) and torch.equal(b.lam, lam0)
    pot = o.make_potential("eight_schools")
    ost = o.arwmh_init(pot, z0)
    ost = ost._replace(adapt_state=o.ARWMHAdaptState(z0.copy(), np.broadcast_to(scale0.astype(ndt), (Cn, d, d)).copy(),
                                                     np.full(Cn, -0.4, ndt)))
    alphas, accs, zs = [], [], []
    for t in range(T):
        ost, alpha, acc = o.arwmh_step(ost, pot, nrm[t], uni[t], adapt=False)
        alphas.append(alpha); accs.append(acc); zs.append(ost.z.copy())
    accs, zs = np.stack(accs), np.stack(zs)
    same = (_np(raw["accept"]).astype(bool) == accs).all(axis=0)
    assert same.mean() >= (1.0 if prec == "f64" else 0.95), same.mean()
    zg = _np(raw["z"]).transpose(0, 2, 1)
    err = (np.abs(zg - zs) / (1 + np.abs(zs))).max(axis=(0, 2))[same]
    assert err.max() < (tol if prec == "f64" else 10 * tol) and np.quantile(err, 0.99) < tol
    np.testing.assert_allclose(_np(b.macc)[same], np.stack(alphas).mean(0)[same], rtol=1e-4, atol=1e-5)


def test_pooled_generic_thread_per_chain_matches_oracle_and_adapts():
    """Pooled adaptation on the thread-per-chain kernel (eight_schools): (i) three windows against the float64 pooled
    oracle on shared draws -- the shared log step size only matches if the window's acceptance rate reaches the update;
    (ii) a long Philox run settles at the 0.234 target."""
    Cn, d, K = 256, 10, 20
    rng = np.random.default_rng(3)
    s = PooledARWMH(models.eight_schools, num_chains=Cn, pool_every=K, dtype=torch.float64)
    s.init(11)
    z = _np(s.batch.z.t()).copy()
    pot = o.make_potential("eight_schools")
    U = pot(z)
    pool = op.pooled_init(z)
    for w in range(3):
        nrm, uni = rng.normal(size=(K, Cn, d)), rng.random(size=(K, Cn))
        dn = torch.from_numpy(nrm).permute(0, 2, 1).contiguous()
        s.run_window(K, collect=(), draws=(dn, torch.from_numpy(uni)))
        z, U, pool, info = op.pooled_window(pot, z, U, pool, K, w * K, draws=(nrm, uni))
        np.testing.assert_allclose(_np(s.batch.macc), info["mean_accept"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(_np(s.batch.z.t()), z, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(float(s.log_step_size), pool["lam"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(_np(s.loc), pool["loc"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(_np(s.dense_scale()), pool["L"], rtol=1e-8, atol=1e-10)
    s = PooledARWMH(models.eight_schools, num_chains=4096, pool_every=50)
    s.init(12)
    s.run(50 * 400, thinning=50, collect=())
    accs = []
    for _ in range(20):
        s.run_window(50, collect=())
        accs.append(float(s.batch.macc.mean()))
    assert abs(np.mean(accs) - 0.234) < 0.03, np.mean(accs)
    assert -3.0 < float(s.log_step_size) < 1.0


def test_pooled_generic_block_kernel_adapts_on_diamonds():
    """The same through the CTA-per-chain kernel (impl = block): before the fix the frozen kernel never stored its
    acceptance rate and the shared step size shrank monotonically."""
    data = models.synthetic_diamonds(n=5000, k=25, seed=0)
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    Cn, d = 512, 26
    q0 = mode[None] + 0.01 * np.random.default_rng(6).normal(size=(Cn, d))
    s = PooledARWMH(models.diamonds, num_chains=Cn, pool_every=50, init_strategy=am.init_to_value(torch.from_numpy(q0)),
                    impl=_lib.IMPL_BLOCK)
    s.init(9, model_kwargs=data)
    s.scale.mul_(0.01)
    s.cov.mul_(1e-4)
    s.run(50 * 80, thinning=50, collect=())
    accs = []
    for _ in range(10):
        s.run_window(50, collect=())
        accs.append(float(s.batch.macc.mean()))
    assert abs(np.mean(accs) - 0.234) < 0.06, np.mean(accs)


def test_sample_Pnx_diamonds_runs_on_tensor_cores():
    """ADVICE (round 1): frozen many-chain diamonds sample_Pnx must reach the tcgen05 shared-state kernel, not one CTA per
    chain.  Same Philox streams through both implementations: identical decisions for almost every chain."""
    data = models.synthetic_diamonds(n=5000, k=25, seed=0)
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    d = 26
    pts = torch.from_numpy(mode[None] + 0.003 * np.random.default_rng(1).normal(size=(8, d))).float()
    ast = am.ARWMHAdaptState(torch.from_numpy(mode).float(), torch.eye(d) * 0.002, torch.tensor(0.0))
    outs = {}
    for impl in (_lib.IMPL_AUTO, _lib.IMPL_BLOCK):
        s = am.ARWMH(models.diamonds)
        s.init(0, 0, None, model_kwargs=data)
        s.impl = impl
        outs[impl] = s.sample_Pnx(21, pts, ast, n=8, n_samples=256)  # 2048 chains > 2 per SM
    a, b = outs[_lib.IMPL_AUTO], outs[_lib.IMPL_BLOCK]
    assert a.shape == (8, 256, d)
    moved = (a - pts[:, None].to(a.device)).abs().amax(-1) > 0
    assert 0.3 < float(moved.float().mean()) <= 1.0
    close = ((a - b).abs().amax(-1) < 2e-5).float().mean()
    assert float(close) > 0.95, float(close)


def test_torch_custom_ops_wrap_the_c_abi():
    """`amcmc::arwmh_run` / `amcmc::logdensity` registered as torch operators over the same C entry points."""
    import adaptive_mcmc_b200.torch_ops  # noqa: F401  (registers torch.ops.amcmc.*)

    Cn, T = 512, 200
    s = am.ARWMH(models.eight_schools, num_chains=Cn)
    st = s.init(6, num_warmup=50, init_params=None)
    b1 = am.ChainBatch.from_state(s.potential, st)
    b2 = am.ChainBatch.from_state(s.potential, st)
    raw = s.run_batch(b1, T, thinning=10)
    h = s.potential.handle.value if hasattr(s.potential.handle, "value") else int(s.potential.handle)
    U = torch.ops.amcmc.logdensity(h, b2.z)
    torch.testing.assert_close(U, b2.pe, rtol=1e-6, atol=1e-5)
    oz, ope = torch.ops.amcmc.arwmh_run(h, b2.z, b2.pe, b2.macc, b2.loc, b2.scale, b2.lam, b2.asc, 0, T, 10, 0, 50, 2 / 3, 0.234, 1e-6,
                                        True, b2.seed, b2.chain_offset, _lib.KERNEL_ARWMH, _lib.IMPL_AUTO)
    assert torch.equal(oz, raw["z"]) and torch.equal(ope, raw["potential_energy"])
    assert torch.equal(b1.z, b2.z) and torch.equal(b1.scale, b2.scale) and torch.equal(b1.lam, b2.lam)


# ---- thread-per-chain kernel behind the per-SM work queue (arwmh_small_balanced_kernel) ---------------------------------

@pytest.mark.parametrize("model,Cn,T,thin,cs", [("eight_schools", 4736 * 3 + 17, 333, 7, 0), ("eight_schools", 1000, 1200, 50, 100),
                                              ("kidiq", 9000, 257, 1, 3), ("eight_schools", 33, 40, 40, 0)])
@pytest.mark.parametrize("adapt", [True, False])
def test_balanced_register_kernel_is_bit_identical_to_the_plain_one(model, Cn, T, thin, cs, adapt):
    """impl = 4 (16 worker warps per SM pop (chain group, segment) items; registers parked raw in shared memory between
    segments) against impl = 1 (a warp owns its chains for the launch): same Philox stream, same arithmetic, so every output
    -- samples, energies, accept decisions, final state incl. as_change -- must be EQUAL, for ragged chain counts (last group
    partial, last CTAs one group short), segment boundaries off the collection points, and the frozen template."""
    pot_model = getattr(models, model)
    outs = []
    for impl in (_lib.IMPL_REGISTER, _lib.IMPL_REGISTER_BALANCED):
        s = am.ARWMH(pot_model, num_chains=Cn)
        s.impl = impl
        kw = dict(model_kwargs=models.synthetic_kidiq()) if model == "kidiq" else {}
        st = s.init(11, num_warmup=100, init_params=None, **kw)
        b = s._batch_from_state(st)
        if not adapt:   # a non-trivial frozen state: adapt for a while first (plain kernel), then freeze
            s.impl = _lib.IMPL_REGISTER
            s.run_batch(b, 150, collect=())
            s.impl = impl
        raw = s.run_batch(b, T, thinning=thin, collect_start=cs, record_accept=True, adapt=adapt)
        outs.append({**{k: v.clone() for k, v in raw.items()}, **{f: getattr(b, f).clone() for f in b._FIELDS}})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("Cn,T,thin,cs,adapt", [(4736 * 2 + 5, 150, 7, 3, True), (1500, 90, 30, 0, False)])
def test_balanced_kernel_runs_asss_bit_identically(Cn, T, thin, cs, adapt):
    """the slice sampler through the same work queue (asss_chain_range): equal outputs for impl 1 and 4"""
    outs = []
    for impl in (_lib.IMPL_REGISTER, _lib.IMPL_REGISTER_BALANCED):
        s = am.ASSS(models.eight_schools, num_chains=Cn)
        s.impl = impl
        b = s._batch_from_state(s.init(4, num_warmup=40, init_params=None))
        if not adapt:
            s.impl = _lib.IMPL_REGISTER
            s.run_batch(b, 60, collect=())
            s.impl = impl
        raw = s.run_batch(b, T, thinning=thin, collect_start=cs, adapt=adapt)
        outs.append({**{k: v.clone() for k, v in raw.items()}, **{f: getattr(b, f).clone() for f in b._FIELDS}})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_balanced_kernel_is_the_default_at_the_headline_size_and_refuses_what_does_not_fit():
    s = am.ARWMH(models.eight_schools, num_chains=64, dtype=torch.float64)
    s.impl = _lib.IMPL_REGISTER_BALANCED
    st = s.init(0, num_warmup=0, init_params=None)
    with pytest.raises(RuntimeError, match="balanced"):
        s.run_batch(s._batch_from_state(st), 10)
    # 65,536 chains: auto = balanced (13.8 groups per SM); equal to the forced plain kernel
    res = []
    for impl in (_lib.IMPL_AUTO, _lib.IMPL_REGISTER):
        s = am.ARWMH(models.eight_schools, num_chains=65536)
        s.impl = impl
        b = s._batch_from_state(s.init(3, num_warmup=50, init_params=None))
        raw = s.run_batch(b, 300, thinning=50)
        res.append((raw["z"].clone(), b.scale.clone(), b.asc.clone()))
    for x, y in zip(*res):
        assert torch.equal(x, y)


@pytest.mark.parametrize("model,Cn,T,thin,cs", [("eight_schools", 1, 130, 7, 3), ("eight_schools", 100, 257, 50, 0), ("kidiq", 1000, 90, 1, 0),
                                              ("eight_schools", 4, 64, 64, 0)])
@pytest.mark.parametrize("adapt", [True, False])
def test_two_warp_few_chain_kernel_is_bit_identical(model, Cn, T, thin, cs, adapt, monkeypatch):
    """few chains: a producer warp generates the draws of step t + 1 into a shared-memory ring while the chain's warp runs step t
    (arwmh_small_duo_kernel).  Same Philox draws and arithmetic as the one-warp kernel: every output must be EQUAL."""
    pot_model = getattr(models, model)
    outs = []
    for duo in ("0", "1"):
        monkeypatch.setenv("AMCMC_SMALL_DUO", duo)
        s = am.ARWMH(pot_model, num_chains=Cn)
        kw = dict(model_kwargs=models.synthetic_kidiq()) if model == "kidiq" else {}
        b = s._batch_from_state(s.init(11, num_warmup=40, init_params=None, **kw))
        if not adapt:
            s.run_batch(b, 50, collect=())
        raw = s.run_batch(b, T, thinning=thin, collect_start=cs, record_accept=True, adapt=adapt)
        outs.append({**{k: v.clone() for k, v in raw.items()}, **{f: getattr(b, f).clone() for f in b._FIELDS}})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("Cn,T,thin,adapt", [(3, 120, 7, True), (500, 90, 30, True), (64, 60, 1, False)])
def test_two_warp_few_chain_asss_is_bit_identical(Cn, T, thin, adapt, monkeypatch):
    """the slice sampler with its head draws produced one step ahead by a second warp (asss_small_duo_kernel)"""
    outs = []
    for duo in ("0", "1"):
        monkeypatch.setenv("AMCMC_SMALL_DUO", duo)
        s = am.ASSS(models.eight_schools, num_chains=Cn)
        b = s._batch_from_state(s.init(4, num_warmup=30, init_params=None))
        if not adapt:
            s.run_batch(b, 40, collect=())
        raw = s.run_batch(b, T, thinning=thin, adapt=adapt)
        outs.append({**{k: v.clone() for k, v in raw.items()}, **{f: getattr(b, f).clone() for f in b._FIELDS}})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
