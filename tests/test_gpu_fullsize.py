"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run
65,536 chains x 10^4 steps in seconds): determinism, shard-independence of every chain's trajectory
(the multi-GPU contract: RNG streams are keyed by the GLOBAL chain id), oracle agreement on a random
subset of chains, and invariants of the sampler (acceptance -> 0.234, energies consistent with positions)."""
import numpy as np
import pytest
import torch

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import _lib, models
from adaptive_mcmc_b200.parallel import PooledARWMH
from oracle import arwmh_numpy as o
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _run_es(C, offset, T, seed=0, thinning=50):
    s = am.ARWMH(models.eight_schools, num_chains=C, chain_offset=offset)
    st = s.init(seed, num_warmup=T // 4, init_params=None)
    b = am.ChainBatch.from_state(s.potential, st, copy=False)
    raw = s.run_batch(b, T, thinning=thinning)
    return b, raw


def test_eight_schools_65536_chains_properties():
    C, T = 65536, 4000
    b1, r1 = _run_es(C, 0, T)
    b2, r2 = _run_es(C, 0, T)
    # (1) bit-level determinism of the fused launch
    assert torch.equal(b1.z, b2.z) and torch.equal(b1.scale, b2.scale) and torch.equal(r1["z"], r2["z"])
    # (2) shard independence: chains [20000, 20064) run alone (as another GPU would) are bit-identical
    b3, r3 = _run_es(64, 20000, T)
    assert torch.equal(b3.z, b1.z[:, 20000:20064]) and torch.equal(b3.scale, b1.scale[:, 20000:20064])
    assert torch.equal(r3["potential_energy"], r1["potential_energy"][:, 20000:20064])
    # (3) stored energies are the potential of the stored positions (fp64 oracle on a random subset)
    idx = torch.randperm(C, generator=torch.Generator().manual_seed(0))[:512].to(b1.z.device)
    zsub = r1["z"][-1][:, idx].t().double().cpu().numpy()
    np.testing.assert_allclose(r1["potential_energy"][-1][idx].cpu().numpy(), o.potential_eight_schools(zsub), rtol=2e-5)
    # (4) oracle agreement of whole trajectories for a few chains of the big batch (Philox stream, fp32 horizon)
    Ts = 100
    s = am.ARWMH(models.eight_schools, num_chains=C)
    st = s.init(3, num_warmup=0, init_params=None)
    bb = am.ChainBatch.from_state(s.potential, st, copy=False)
    raw = s.run_batch(bb, Ts, record_accept=True)
    pick = np.array([0, 1, 31, 32, 4095, 32768, 65535])
    q0 = co.init_uniform(3, C, 10, dt=np.float32)[pick]
    ost = o.arwmh_init(o.make_potential("eight_schools"), q0)
    for k, c in enumerate(pick):
        one = o.ARWMHState(0, ost.z[k:k + 1], ost.potential_energy[k:k + 1], ost.mean_accept_prob[k:k + 1],
                           o.ARWMHAdaptState(ost.adapt_state.loc[k:k + 1], ost.adapt_state.scale[k:k + 1],
                                             ost.adapt_state.log_step_size[k:k + 1]), ost.as_change[k:k + 1], 0)
        ol, oc = co.arwmh_run(one, "eight_schools", Ts, seed=3, chain_offset=int(c), record_accept=True)
        acc_g = raw["accept"][:, c].cpu().numpy().astype(bool)
        if (acc_g == oc["accepts"][:, 0]).all():
            zg = raw["z"][:, :, c].cpu().numpy()
            assert (np.abs(zg - oc["z"][:, 0]) / (1 + np.abs(oc["z"][:, 0]))).max() < 1e-2
    # (5) sampler invariants on the whole population
    acc = float(b1.macc.mean())
    assert 0.2 < acc < 0.28, acc
    assert torch.isfinite(b1.scale).all() and float(r1["potential_energy"].min()) > 40.05


def test_diamonds_tc_65536_chains_properties():
    data = models.synthetic_diamonds(n=5000, k=25, seed=0)
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    C = 65536
    q0 = mode[None] + 0.01 * np.random.default_rng(0).normal(size=(C, 26))

    def run(Cn, offset, q):
        s = PooledARWMH(models.diamonds, num_chains=Cn, pool_every=50, chain_offset=offset,
                        init_strategy=am.init_to_value(torch.from_numpy(q)), impl=_lib.IMPL_TENSOR)
        s.init(1, model_kwargs=data)
        s.scale.mul_(0.003)
        raw = s.run_window(50, thinning=50, adapt=False)
        return s, raw

    s1, r1 = run(C, 0, q0)
    s2, r2 = run(C, 0, q0)
    assert torch.equal(s1.batch.z, s2.batch.z) and torch.equal(s1.batch.pe, s2.batch.pe)  # determinism
    # shard independence needs the same shared proposal: loc is the global mean, so give the shard the same loc
    s3 = PooledARWMH(models.diamonds, num_chains=256, pool_every=50, chain_offset=30000,
                     init_strategy=am.init_to_value(torch.from_numpy(q0[30000:30256])), impl=_lib.IMPL_TENSOR)
    s3.init(1, model_kwargs=data)
    s3.scale.mul_(0.003)
    s3.loc.copy_(torch.from_numpy(q0.mean(0)).to(s3.loc))
    s4 = PooledARWMH(models.diamonds, num_chains=C, pool_every=50, init_strategy=am.init_to_value(torch.from_numpy(q0)),
                     impl=_lib.IMPL_TENSOR)
    s4.init(1, model_kwargs=data)
    s4.scale.mul_(0.003)
    s4.loc.copy_(s3.loc)
    s3.run_window(50, adapt=False, collect=())
    s4.run_window(50, adapt=False, collect=())
    assert torch.equal(s3.batch.z, s4.batch.z[:, 30000:30256]) and torch.equal(s3.batch.pe, s4.batch.pe[30000:30256])
    # energies vs the fp64 oracle on a subset; acceptance in a sane band
    idx = np.random.default_rng(1).choice(C, 256, replace=False)
    zs = s1.batch.z[:, torch.from_numpy(idx).to(s1.batch.z.device)].t().double().cpu().numpy()
    U = o.potential_diamonds(zs, X, Y)
    assert np.abs(s1.batch.pe.cpu().numpy()[idx] - U).max() < 1e-2
    assert 0.05 < float(s1.batch.macc.mean()) < 0.8


@pytest.mark.parametrize("C", [65536, 131072])
def test_diamonds_tc_adaptive_full_size_properties(C):
    """`diamonds_tc_adapt_kernel` (per-chain adaptation = the reference's ARWMH.sample, arwmh.py:140-207) at the chain
    counts of BASELINE.json configs[2]-style runs (65,536 on one GPU) and of a configs[3] slice (131,072 per GPU):
    determinism, shard independence under `chain_offset` (the multi-GPU contract), energies against the float64 oracle on a
    random subset, acceptance band, finite adaptation state.  131,072 chains = 1024 groups = two rounds per CTA."""
    data = models.synthetic_diamonds(n=5000, k=25, seed=0)
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    q0 = mode[None] + 0.004 * np.random.default_rng(0).normal(size=(C, 26))
    T = 300  # more than one 256-step segment: the per-chain GEMM reference points move once

    def run(Cn, offset, q):
        s = am.ARWMH(models.diamonds, num_chains=Cn, chain_offset=offset, init_strategy=am.init_to_value(torch.from_numpy(q)))
        s.impl = _lib.IMPL_TENSOR
        st = s.init(2, num_warmup=0, init_params=None, model_kwargs=data)
        b = am.ChainBatch.from_state(s.potential, st, copy=False)
        b.set_dense_scale(torch.eye(26) * 0.002)
        raw = s.run_batch(b, T, thinning=T, collect=("z", "potential_energy"))
        return b, raw

    b1, r1 = run(C, 0, q0)
    b2, r2 = run(C, 0, q0)
    for f in ("z", "pe", "loc", "scale", "lam", "macc"):
        assert torch.equal(getattr(b1, f), getattr(b2, f)), f
    lo = C - 40000  # a shard that straddles group and round boundaries, run alone with its global chain ids
    b3, r3 = run(384, lo, q0[lo:lo + 384])
    assert torch.equal(b3.z, b1.z[:, lo:lo + 384]) and torch.equal(b3.scale, b1.scale[:, lo:lo + 384])
    assert torch.equal(b3.lam, b1.lam[lo:lo + 384]) and torch.equal(r3["potential_energy"], r1["potential_energy"][:, lo:lo + 384])
    idx = np.random.default_rng(1).choice(C, 384, replace=False)
    ti = torch.from_numpy(idx).to(b1.z.device)
    U = o.potential_diamonds(b1.z[:, ti].t().double().cpu().numpy(), X, Y)
    assert np.abs(b1.pe.cpu().numpy()[idx] - U).max() < 2e-2
    np.testing.assert_allclose(r1["potential_energy"][-1].cpu().numpy(), b1.pe.cpu().numpy())
    acc = float(b1.macc.mean())
    assert 0.1 < acc < 0.5, acc
    assert torch.isfinite(b1.scale).all() and torch.isfinite(b1.lam).all() and torch.isfinite(b1.loc).all()
    assert int(b1.i) == T


def test_ram_gaussian_16384_chains_properties():
    """BASELINE.json configs[4] at full size: 16,384 robust-adaptive-Metropolis chains on the d = 200 AR(1) Gaussian."""
    d, C, T = 200, 16384, 60
    P = models.ar1_precision_chol(d, 0.9)

    def run(Cn, offset):
        s = am.RAM(models.gaussian, num_chains=Cn, chain_offset=offset, init_strategy=am.init_to_value(torch.zeros(Cn, d)))
        st = s.init(4, num_warmup=0, init_params=None, model_kwargs=dict(prec_chol=P))
        b = am.ChainBatch.from_state(s.potential, st, copy=False)
        b.scale.mul_(2.38 / d**0.5 * 0.3)
        raw = s.run_batch(b, T, thinning=T, collect=("z", "potential_energy"))
        return b, raw

    b1, r1 = run(C, 0)
    b2, _ = run(C, 0)
    assert torch.equal(b1.z, b2.z) and torch.equal(b1.scale, b2.scale) and torch.equal(b1.pe, b2.pe)
    b3, _ = run(96, 9000)
    assert torch.equal(b3.z, b1.z[:, 9000:9096]) and torch.equal(b3.scale, b1.scale[:, 9000:9096])
    idx = np.random.default_rng(2).choice(C, 128, replace=False)
    z = b1.z[:, torch.from_numpy(idx).to(b1.z.device)].t().double().cpu().numpy()
    U = 0.5 * ((z @ np.tril(np.asarray(P, np.float64))) ** 2).sum(1)
    np.testing.assert_allclose(b1.pe.cpu().numpy()[idx], U, rtol=2e-4, atol=2e-4)
    acc = float(b1.macc.mean())
    assert 0.1 < acc < 0.6, acc
    assert torch.isfinite(b1.scale).all() and (b1.scale.abs().amax() < 10)
