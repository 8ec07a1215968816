"""The reference's quality experiment at its own scale (run_eight_schools_wasserstein.py:63 + eval_eight_schools.py):
100 seeds x (50k warm-up + 500k samples, thinning 50) as 100 chains of one launch, then rmse_means / wasserstein /
mmd per seed against 10^4 independent posterior draws.  The recorded table of the reference
(posteriordb_eight-schools.ipynb:L2038, tests/golden/reference_pins.json) is reproduced in distribution; its y are
the posteriordb Stan draws, ours are independent chains' end states, so the bands are a few recorded sds wide."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200.utils import evaluation as ev

pytestmark = pytest.mark.gpu

RECORDED = {"rmse_means": (0.0745, 0.0177), "wasserstein": (1.6865, 0.0028), "mmd": (0.01569, 0.00112)}  # arwm, 100 seeds


def test_eight_schools_quality_table():
    from eval_eight_schools import reference_draws, unconstrained

    y = reference_draws()
    mcmc = am.MCMC(am.ARWMH(am.models.eight_schools), num_warmup=50_000, num_samples=500_000, thinning=50, num_chains=100)
    mcmc.run(0)
    x = unconstrained(mcmc.get_samples(group_by_chain=True))
    assert x.shape == (100, 10000, 10)
    rmse = np.array([ev.pth_moment_rmse(x[k].contiguous(), y, p=1) for k in range(100)])
    mmd = np.array([ev.mmd_heuristic(x[k].contiguous(), y) for k in range(100)])
    # per-seed spread like the recorded one, mean inside +-2 recorded sds (rmse) / the NUTS..ARWM band (mmd)
    assert abs(rmse.mean() - RECORDED["rmse_means"][0]) < 2 * RECORDED["rmse_means"][1], rmse.mean()
    assert 0.5 * RECORDED["rmse_means"][1] < rmse.std() < 2.5 * RECORDED["rmse_means"][1], rmse.std()
    assert 0.0125 < mmd.mean() < 0.0180 and mmd.std() < 2 * RECORDED["mmd"][1], (mmd.mean(), mmd.std())
    # the 10^4 x 10^4 assignment for one seed (the reference records 20.7 s per call for the host solver)
    w = ev.wasserstein_dist11_p(x[0].contiguous(), y)
    assert abs(w - RECORDED["wasserstein"][0]) < 0.04, w
    # sanity of the yardstick itself: y against a second independent set of draws sits at the same noise floor
    y2 = reference_draws(seed=999)
    assert ev.mmd_heuristic(y2, y) < 0.0165 and ev.pth_moment_rmse(y2, y, p=1) < 0.15
