"""GPU samplers against the posteriordb reference draws the reference repo ships
(`python/mcmc_runs/diamonds-example-references.pkl`, moments in tests/golden/reference_pins.json) on the
diamonds-equivalent data set recovered from them (tests/test_diamonds_pin.py explains the recovery and what it leaves
unfitted).  Coordinates in the order of python/scripts/eval_diamonds.py:78-87: [Intercept, b[0..23], log sigma].

Three samplers, the same bar -- pooled posterior mean within a few MCSE of the 10,000 reference draws and posterior sd
within a few per cent in every coordinate, including the unfitted log sigma:

* the exact CUDA-core block kernel (per-chain adaptation, the reference's algorithm as is),
* the tcgen05 kernel with per-chain adaptation (the same algorithm, likelihood on the tensor cores),
* the tcgen05 shared-state kernel driven by pooled adaptation, then FROZEN (an exactly invariant kernel).

The reference's own algorithm carries a finite-adaptation under-dispersion (with gamma = n^-2/3 the
proposal covariance is an average over the last ~1/gamma steps), so the per-chain-adaptive runs get the wider sd band:
the float64 C ORACLE run on this data set with the block-kernel configuration below (48 chains x 120,000 steps, second
half kept) gives sd ratios 0.871 ... 0.928 against the reference draws and means within 5.6 MCSE -- the CUDA kernels
must land in the same place (measured: block kernel 0.878 ... 1.00), not at 1.
"""
import json
import os

import numpy as np
import pytest
import torch

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import _lib, models
from adaptive_mcmc_b200.parallel import PooledARWMH
from oracle import arwmh_numpy as o
from oracle import diamonds_exact as de

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PINS = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_pins.json")))
REF_MEAN = np.array(PINS["diamonds_reference_draws"]["mean"])
REF_SD = np.array(PINS["diamonds_reference_draws"]["std"])
MCSE = REF_SD / np.sqrt(PINS["diamonds_reference_draws"]["n"])


@pytest.fixture(scope="module")
def pinned_data():
    r = PINS["diamonds_recovered_stats"]
    return de.dataset_from_stats(dict(n=r["n"], G=np.array(r["G"]), h=np.array(r["h"]), yy=r["yy"]), seed=0)


def _report(tag, z):
    """z: [n, 26] pooled draws.  Returns (max |mean err| in reference MCSE, sd ratios)."""
    zm = np.abs(z.mean(0) - REF_MEAN) / MCSE
    ratio = z.std(0) / REF_SD
    print(f"{tag}: mean err max {zm.max():.2f} MCSE (coord {zm.argmax()}), sd ratio [{ratio.min():.3f}, {ratio.max():.3f}], "
          f"log-sigma sd ratio {ratio[25]:.3f}")
    return zm, ratio


def test_potential_on_pinned_data_matches_oracle(pinned_data):
    q = REF_MEAN[None] + REF_SD[None] * np.random.default_rng(0).normal(size=(64, 26))
    want = o.make_potential("diamonds", X=pinned_data["X"], Y=pinned_data["Y"])(q)
    for dt, tol in ((torch.float64, 1e-10), (torch.float32, 3e-5)):  # points far off the posterior: U up to 1e5
        pot = models.diamonds.bind(dtype=dt, **pinned_data)
        got = pot(torch.from_numpy(q).to(dt)).double().cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=tol)


def _start(C, seed):
    return REF_MEAN[None] + 0.5 * REF_SD[None] * np.random.default_rng(seed).normal(size=(C, 26))


@pytest.mark.parametrize("impl,C,steps", [(_lib.IMPL_BLOCK, 296, 120_000), (_lib.IMPL_TENSOR, 4096, 300_000)])
def test_adaptive_samplers_reproduce_reference_draws(impl, C, steps, pinned_data):
    """ARWMH.sample (arwmh.py:140-207) with per-chain adaptation, lr_decay = 2/3, started at the posterior mean with a
    small isotropic factor (the reference's identity start rejects every proposal on this posterior: its own runs use a
    10^6-step warm-up)."""
    s = am.ARWMH(models.diamonds, num_chains=C, init_strategy=am.init_to_value(torch.from_numpy(_start(C, 1))))
    s.impl = impl
    st = s.init(5, num_warmup=0, init_params=None, model_kwargs=pinned_data)
    b = am.ChainBatch.from_state(s.potential, st)
    b.set_dense_scale(torch.eye(26) * 0.003)
    burn = steps // 2
    raw = s.run_batch(b, steps, thinning=(steps - burn) // 50, collect_start=burn, collect=("z",))
    z = raw["z"].double().permute(0, 2, 1).reshape(-1, 26).cpu().numpy()
    assert np.isfinite(z).all()
    zm, ratio = _report(f"adaptive impl={impl} C={C}", z)
    acc = float(b.macc.mean())
    assert abs(acc - 0.234) < 0.03, acc
    # pooled over C chains x 50 states the Monte-Carlo error of our mean is far below one reference MCSE; what is left
    # is the reference draws' own error (1 MCSE = 1 sd of that) and the sampler's bias
    assert zm.max() < 8.0, zm
    assert 0.85 < ratio.min() and ratio.max() < 1.06, ratio


def test_pooled_then_frozen_tensor_core_sampler_reproduces_reference_draws(pinned_data):
    """Pooled adaptation on the tcgen05 shared-state kernel, then the adaptation is switched off: the frozen random-walk
    kernel is exactly invariant, so the pooled sample has no adaptation bias and must match the draws tightly."""
    C = 8192
    s = PooledARWMH(models.diamonds, num_chains=C, pool_every=100, init_strategy=am.init_to_value(torch.from_numpy(_start(C, 2))))
    s.init(9, model_kwargs=pinned_data)
    s.scale.mul_(0.003)
    s.cov.mul_(0.003**2)
    s.run(100 * 400, thinning=100, collect=())           # adapt: 400 windows
    zs = []
    for w in range(600):                                   # frozen: 60,000 steps, one state per 1,500
        raw = s.run_window(100, thinning=100, collect=("z",) if w % 15 == 14 else (), adapt=False)
        if "z" in raw:
            zs.append(raw["z"][-1].t().double().cpu().numpy())
    z = np.concatenate(zs, 0)
    zm, ratio = _report("pooled+frozen tcgen05", z)
    assert zm.max() < 5.0, zm
    assert 0.97 < ratio.min() and ratio.max() < 1.03, ratio
