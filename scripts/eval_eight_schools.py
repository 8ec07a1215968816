"""Quality evaluation of eight_schools runs -- the GPU counterpart of the reference's
python/scripts/run_eight_schools_wasserstein.py + eval_eight_schools.py:

  * 100 seeds = 100 independent ARWMH chains in ONE launch, 50k warm-up + 500k samples, thinning 50
    (run_eight_schools_wasserstein.py:63) -> 10^4 kept draws per seed;
  * per seed: rmse_means = pth_moment_rmse(x, y, p=1), wasserstein = wasserstein_dist11_p(x, y),
    mmd = mmd_heuristic(x, y) in the unconstrained coordinates [mu, log tau, theta_base[8]]
    (eval_eight_schools.py:48-56, 66-80).

The reference's y are the posteriordb reference draws (10 Stan chains x 1000), which need the posteriordb checkout.
Here y are 10^4 EXACT independent posterior draws (`exact_draws`: the model is conditionally conjugate, so tau is drawn
from its one-dimensional marginal by quadrature and mu, theta | tau from their Gaussian conditionals), i.e. the same
yardstick in distribution.  Recorded table (posteriordb_eight-schools.ipynb:L2038):
    arwm  rmse_means 0.0745 +- 0.0177   wasserstein 1.6865 +- 0.0028   mmd 0.01569 +- 0.00112
    asss  rmse_means 0.0607             wasserstein 1.7009             mmd 0.01478
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_mcmc_b200 as am  # noqa: E402
from adaptive_mcmc_b200.utils import evaluation as ev  # noqa: E402


def unconstrained(samples):
    """[C, S] / [C, S, 8] site tensors -> [C, S, 10] in ravel order (mu, log tau, theta_base)."""
    return torch.cat([samples["mu"][..., None], torch.log(samples["tau"])[..., None], samples["theta_base"]], dim=-1)


Y_OBS = np.array([28, 8, -3, 7, -1, 1, 18, 12.0])
SIGMA = np.array([15, 10, 16, 11, 9, 11, 10, 18.0])


def _tau_marginal(t):
    """log p(t = log tau | y) up to a constant, and mean / precision of mu | tau, y  (mu and theta integrated out:
    y_j | mu, tau ~ N(mu, sigma_j^2 + tau^2), mu ~ N(0, 5^2), tau ~ HalfCauchy(5))."""
    tau = np.exp(t)
    V = SIGMA[None, :] ** 2 + tau[:, None] ** 2
    prec = 1 / 25 + (1 / V).sum(1)
    mean = (Y_OBS[None, :] / V).sum(1) / prec
    logm = -0.5 * np.log(V).sum(1) - 0.5 * (Y_OBS[None, :] ** 2 / V).sum(1) + 0.5 * mean**2 * prec - 0.5 * np.log(prec)
    return logm + np.log(2 / (np.pi * 5 * (1 + (tau / 5) ** 2))) + t, mean, prec


def exact_moments():
    t = np.linspace(-30, 8, 800001)
    lp, mean, prec = _tau_marginal(t)
    w = np.exp(lp - lp.max())
    w /= w.sum()
    et = (w * t).sum()
    emu = (w * mean).sum()
    return {"log_tau": (et, np.sqrt((w * (t - et) ** 2).sum())), "mu": (emu, np.sqrt((w * (1 / prec + mean**2)).sum() - emu**2)),
            "tau_mean": (w * np.exp(t)).sum(), "P(log_tau<-2)": w[t < -2].sum()}


def exact_draws(n=10000, seed=0, device="cuda"):
    """n independent draws from the eight_schools posterior in the unconstrained coordinates [mu, log tau, theta_base]."""
    rng = np.random.default_rng(seed)
    t = np.linspace(-30, 8, 800001)
    lp, _, _ = _tau_marginal(t)
    cdf = np.cumsum(np.exp(lp - lp.max()))
    cdf /= cdf[-1]
    ts = np.interp(rng.random(n), cdf, t)
    _, mean, prec = _tau_marginal(ts)
    mu = mean + rng.normal(size=n) / np.sqrt(prec)
    tau = np.exp(ts)
    # theta_base_j | mu, tau, y: prior N(0,1), y_j ~ N(mu + tau eta_j, sigma_j^2)
    pj = 1 + tau[:, None] ** 2 / SIGMA[None, :] ** 2
    mj = tau[:, None] * (Y_OBS[None, :] - mu[:, None]) / SIGMA[None, :] ** 2 / pj
    eta = mj + rng.normal(size=(n, 8)) / np.sqrt(pj)
    out = np.concatenate([mu[:, None], ts[:, None], eta], axis=1).astype(np.float32)
    return torch.from_numpy(out).to(device)


def reference_draws(n=10000, steps=200000, seed=12345):
    """Alternative yardstick: the end states of n independent adaptive chains (carries the sampler's own tail bias)."""
    s = am.ARWMH(am.models.eight_schools, num_chains=n)
    st = s.init(seed, num_warmup=steps // 2, init_params=None)
    _, last = s.run(st, steps, collect=())
    z = last.z
    return torch.cat([z["mu"][:, None], z["tau"][:, None], z["theta_base"]], dim=1)  # unconstrained already


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="rwm", choices=["rwm", "sss"], help="ARWMH or ASSS (run_eight_schools_wasserstein.py:58-66)")
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--num-warmup", type=int, default=None)
    ap.add_argument("--num-samples", type=int, default=None)
    ap.add_argument("--thinning", type=int, default=None)
    ap.add_argument("--wasserstein-seeds", type=int, default=2, help="seeds that also get the 10^4 x 10^4 assignment (20 s of host time each)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dflt = dict(rwm=(50_000, 500_000, 50), sss=(25_000, 250_000, 25))[a.kernel]  # the reference's sample_params
    a.num_warmup = dflt[0] if a.num_warmup is None else a.num_warmup
    a.num_samples = dflt[1] if a.num_samples is None else a.num_samples
    a.thinning = dflt[2] if a.thinning is None else a.thinning
    t0 = time.time()
    y = exact_draws()
    torch.cuda.synchronize()
    t1 = time.time()
    sampler = am.ARWMH(am.models.eight_schools) if a.kernel == "rwm" else am.ASSS(am.models.eight_schools)
    mcmc = am.MCMC(sampler, num_warmup=a.num_warmup, num_samples=a.num_samples, thinning=a.thinning,
                   num_chains=a.seeds)
    mcmc.run(0)
    x_all = unconstrained(mcmc.get_samples(group_by_chain=True))  # [seeds, S, 10]
    torch.cuda.synchronize()
    t2 = time.time()
    rows = []
    for k in range(a.seeds):
        x = x_all[k].contiguous()
        row = {"rng_seed": k, "rmse_means": ev.pth_moment_rmse(x, y, p=1), "mmd": ev.mmd_heuristic(x, y)}
        if k < a.wasserstein_seeds:
            row["wasserstein"] = ev.wasserstein_dist11_p(x, y)
        rows.append(row)
    t3 = time.time()
    agg = {m: (float(np.mean([r[m] for r in rows if m in r])), float(np.std([r[m] for r in rows if m in r])))
           for m in ("rmse_means", "wasserstein", "mmd") if any(m in r for r in rows)}
    pooled = x_all.reshape(-1, 10).double()
    print(json.dumps({"exact_moments": {k: (list(map(float, v)) if isinstance(v, tuple) else float(v)) for k, v in exact_moments().items()},
                      "pooled_sample": {"mu": [float(pooled[:, 0].mean()), float(pooled[:, 0].std())],
                                        "log_tau": [float(pooled[:, 1].mean()), float(pooled[:, 1].std())],
                                        "P(log_tau<-2)": float((pooled[:, 1] < -2).double().mean())}}))
    print(json.dumps({"agg_mean_std": agg, "seconds": {"reference_draws": t1 - t0, "runs": t2 - t1, "metrics": t3 - t2},
                      "kept_per_seed": int(x_all.shape[1])}))
    if a.out:
        json.dump(rows, open(a.out, "w"))


if __name__ == "__main__":
    main()
