"""The reference's quality experiment at its own scale (run_eight_schools_wasserstein.py:58-66 + eval_eight_schools.py):
100 seeds x (50k warm-up + 500k samples, thinning 50) for ARWMH, 100 x (25k + 250k, thinning 25) for ASSS, each as 100
chains of one launch; then rmse_means / wasserstein / mmd per seed against 10^4 reference draws.  The reference's y are
the posteriordb Stan draws; ours are EXACT independent posterior draws (the model is conditionally conjugate:
scripts/eval_eight_schools.py:exact_draws), the same yardstick in distribution.  The table the reference recorded
(posteriordb_eight-schools.ipynb:L2038, tests/golden/reference_pins.json:quality_tables) is reproduced."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200.utils import evaluation as ev

TAB = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_pins.json")))["quality_tables"]["eight_schools"]


def test_exact_posterior_draws_match_quadrature():
    from eval_eight_schools import exact_draws, exact_moments

    mom = exact_moments()
    assert abs(mom["tau_mean"] - 3.5977) < 1e-3 and abs(mom["log_tau"][0] - 0.80214) < 1e-4
    y = exact_draws(400_000, seed=3, device="cpu").double().numpy()
    se = y.std(0) / np.sqrt(len(y))
    assert abs(y[:, 1].mean() - mom["log_tau"][0]) < 4 * se[1] and abs(y[:, 1].std() - mom["log_tau"][1]) < 0.01
    assert abs(y[:, 0].mean() - mom["mu"][0]) < 4 * se[0] and abs(y[:, 0].std() - mom["mu"][1]) < 0.02
    assert abs((y[:, 1] < -2).mean() - mom["P(log_tau<-2)"]) < 1.5e-3
    # the reference's recorded single-run table (ARWMH): every mean within 0.1 sd of the exact draws
    pins = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_pins.json")))["eight_schools_arwmh_table"]
    cons = np.concatenate([y[:, :1], np.exp(y[:, 1:2]), y[:, 2:]], axis=1)  # tau constrained, as the table reports it
    assert (np.abs(cons.mean(0) - np.array(pins["mean"])) / cons.std(0)).max() < 0.1


@pytest.mark.gpu
def test_eight_schools_quality_table():
    from eval_eight_schools import exact_draws, unconstrained

    y = exact_draws(10000, seed=0)
    res = {}
    for name, sampler, cfg in (("arwm", am.ARWMH(am.models.eight_schools), (50_000, 500_000, 50)),
                               ("asss", am.ASSS(am.models.eight_schools), (25_000, 250_000, 25))):
        mcmc = am.MCMC(sampler, num_warmup=cfg[0], num_samples=cfg[1], thinning=cfg[2], num_chains=100)
        mcmc.run(0)
        x = unconstrained(mcmc.get_samples(group_by_chain=True))
        assert x.shape == (100, 10000, 10)
        rmse = np.array([ev.pth_moment_rmse(x[k].contiguous(), y, p=1) for k in range(100)])
        mmd = np.array([ev.mmd_heuristic(x[k].contiguous(), y) for k in range(100)])
        w = float(np.mean([ev.wasserstein_dist11_p(x[k].contiguous(), y) for k in range(8)]))  # 10^4 x 10^4 assignments on the GPU
        res[name] = (rmse, mmd, w)
    # ---- ARWMH: recorded 0.0745 +- 0.0177 / 1.6865 +- 0.0028 / 0.01569 +- 0.00112 (mean +- sd over 100 seeds)
    rmse, mmd, w = res["arwm"]
    rec = TAB["arwm"]
    assert abs(rmse.mean() - rec["rmse_means"][0]) < 0.007, rmse.mean()          # 3 standard errors of either mean
    assert 0.6 * rec["rmse_means"][1] < rmse.std() < 1.6 * rec["rmse_means"][1], rmse.std()
    assert abs(mmd.mean() - rec["mmd"][0]) < 0.0012, mmd.mean()
    assert 0.5 * rec["mmd"][1] < mmd.std() < 1.6 * rec["mmd"][1], mmd.std()
    assert abs(w - rec["wasserstein"][0]) < 0.012, w
    # ---- ASSS: recorded 0.0607 / 1.7009 / 0.01478
    rmse_s, mmd_s, w_s = res["asss"]
    rec = TAB["asss"]
    assert abs(rmse_s.mean() - rec["rmse_means"][0]) < 0.008, rmse_s.mean()
    assert abs(mmd_s.mean() - rec["mmd"][0]) < 0.0015, mmd_s.mean()
    assert abs(w_s - rec["wasserstein"][0]) < 0.015, w_s
    # ---- and the ordering the reference reports between the two samplers
    assert mmd.mean() > mmd_s.mean() and rmse.mean() > rmse_s.mean() and w < w_s


@pytest.mark.gpu
def test_diamonds_quality_table():
    """The reference's 100-seed quality experiment for diamonds (run_diamonds_wasserstein.py + eval_diamonds.py): 100 seeds as
    100 chains of one launch on the CTA-per-chain kernel, from the reference's own start (q0 ~ U(-2,2)^26, identity factor),
    metrics against the reference's own y -- the posteriordb draws it ships -- on the diamonds-equivalent data set recovered
    from them (tests/test_diamonds_pin.py).  Recorded (posteriordb_diamonds.ipynb:L3635, mean +- sd over 100 seeds):
    rmse_means 0.01566 +- 0.0074, wasserstein 0.12315 +- 0.00126, mmd 0.03310 +- 0.00346.

    Run length: the committed script says 10^6 warm-up + 10^7 samples, thinning 1000 (run_diamonds_wasserstein.py:67); with that
    the GPU run is BETTER than the recorded table in all three metrics (0.0062 / 0.12049 / 0.0151, at the NUTS / ASSS floor the
    reference records: 0.0107 / 0.12183 / 0.0142).  The recorded table is reproduced -- means AND seed-to-seed spreads -- by
    10^6 warm-up + 10^6 samples, thinning 100 (measured: 0.01501 +- 0.00707 / 0.12327 +- 0.00115 / 0.03255 +- 0.00328), which is
    evidently what produced it (the committed script cannot run as it is: SURVEY section 3.1).  That configuration is the test."""
    from eval_diamonds import RECORDED, run

    out, rows = run(seeds=100, w1_seeds=12, num_warmup=1_000_000, num_samples=1_000_000, thinning=100)
    print(json.dumps(out))
    assert abs(out["accept"] - 0.234) < 0.01
    r = out["rmse_means"]
    assert abs(r["mean"] - RECORDED["rmse_means"][0]) < 3 * (RECORDED["rmse_means"][1] + r["sd"]) / 10, r   # 3 SE of both means
    assert 0.6 * RECORDED["rmse_means"][1] < r["sd"] < 1.6 * RECORDED["rmse_means"][1], r
    m = out["mmd"]
    assert abs(m["mean"] - RECORDED["mmd"][0]) < 3 * (RECORDED["mmd"][1] + m["sd"]) / 10, m
    assert 0.6 * RECORDED["mmd"][1] < m["sd"] < 1.6 * RECORDED["mmd"][1], m
    w = out["wasserstein"]
    assert abs(w["mean"] - RECORDED["wasserstein"][0]) < 3 * RECORDED["wasserstein"][1] / np.sqrt(12) + 3 * RECORDED["wasserstein"][1] / 10, w
