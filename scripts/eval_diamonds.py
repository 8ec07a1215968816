"""Quality evaluation of diamonds runs -- the GPU counterpart of the reference's python/scripts/run_diamonds_wasserstein.py +
eval_diamonds.py, against the numbers the reference recorded for its 100-seed experiment
(python/jupyter/posteriordb_diamonds.ipynb:L3635):   arwm  rmse_means 0.01566 +- 0.0074   wasserstein 0.12315 +- 0.00126
mmd 0.03310 +- 0.00346.

  * 100 seeds = 100 independent ARWMH chains in ONE launch from the reference's own start (init_to_uniform: q0 ~ U(-2,2)^26,
    identity factor), 10^6 warm-up + 10^7 samples, thinning 1000 (run_diamonds_wasserstein.py:67) -> 10^4 kept draws per seed;
  * per seed: rmse_means = pth_moment_rmse(x, y, p=1), wasserstein = wasserstein_dist11_p(x, y), mmd = mmd_heuristic(x, y) in the
    coordinates [Intercept, b[24], log sigma] (eval_diamonds.py:78-104);
  * y = the posteriordb reference draws the reference ships (python/mcmc_runs/diamonds-example-references.pkl, fixture
    tests/golden/diamonds_reference_draws.npz) -- the reference's own y;
  * data = the diamonds-equivalent data set recovered from those draws (tests/test_diamonds_pin.py): posteriordb's diamonds.json is
    not in the image, the likelihood depends on the data only through its sufficient statistics.

    python scripts/eval_diamonds.py [--scale 1.0] [--w1-seeds 100]      (--scale 0.1: ten times shorter runs)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adaptive_mcmc_b200 as am  # noqa: E402
from adaptive_mcmc_b200 import models  # noqa: E402
from adaptive_mcmc_b200.utils import evaluation as ev  # noqa: E402

RECORDED = {"rmse_means": (0.01566, 0.0074), "wasserstein": (0.12315, 0.00126), "mmd": (0.03310, 0.00346)}


def pinned_data():
    r = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_pins.json")))["diamonds_recovered_stats"]
    return models.diamonds_from_sufficient_stats(r["n"], r["G"], r["h"], r["yy"], seed=0)


def reference_draws(device="cuda"):
    return torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "diamonds_reference_draws.npz"))["y"]).to(device)


RECORDED_ASSS = {"rmse_means": (0.009635, 0.003629), "wasserstein": (0.121571, 0.000778), "mmd": (0.013965, 0.001473)}


def run(seeds=100, scale=1.0, w1_seeds=100, dtype=torch.float32, num_warmup=None, num_samples=None, thinning=None, kernel="rwm"):
    num_warmup = int(1_000_000 * scale) if num_warmup is None else int(num_warmup)
    num_samples = int(10_000_000 * scale) if num_samples is None else int(num_samples)
    thinning = max(1, int(1000 * scale)) if thinning is None else int(thinning)
    data = pinned_data()
    y = reference_draws()
    # the reference's defaults: lr_decay 2/3, target 0.234; "sss" = the adaptive stereographic slice sampler (python/kernels/asss.py)
    sampler = am.ARWMH(models.diamonds, dtype=dtype) if kernel == "rwm" else am.ASSS(models.diamonds, dtype=dtype)
    rec = RECORDED if kernel == "rwm" else RECORDED_ASSS
    mcmc = am.MCMC(sampler, num_warmup=num_warmup, num_samples=num_samples, thinning=thinning, num_chains=seeds)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mcmc.run(0, **data)
    torch.cuda.synchronize()
    t_sample = time.perf_counter() - t0
    smp = mcmc.get_samples(group_by_chain=True)                          # constrained sites [C, S, ...]
    x = torch.cat([smp["Intercept"][..., None], smp["b"], torch.log(smp["sigma"])[..., None]], dim=-1).float()  # eval_diamonds.py:78-87
    rows = []
    t0 = time.perf_counter()
    for c in range(seeds):
        row = {"rng_seed": c, "rmse_means": ev.pth_moment_rmse(x[c], y, p=1.0), "mmd": ev.mmd_heuristic(x[c], y)}
        if c < w1_seeds:
            row["wasserstein"] = ev.wasserstein_dist11_p(x[c], y)
        rows.append(row)
    torch.cuda.synchronize()
    t_eval = time.perf_counter() - t0
    out = {"seeds": seeds, "num_warmup": num_warmup, "num_samples": num_samples, "thinning": thinning,
           "sampling_s": t_sample, "metrics_s": t_eval, "kernel": kernel,
           "accept": float(mcmc.last_state.mean_accept_prob.mean()) if kernel == "rwm" else None}
    for k in ("rmse_means", "wasserstein", "mmd"):
        v = np.array([r[k] for r in rows if k in r])
        out[k] = {"mean": float(v.mean()), "sd": float(v.std(ddof=1)) if len(v) > 1 else None, "n": int(len(v)),
                  "recorded_mean": rec[k][0], "recorded_sd": rec[k][1]}
    return out, rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--w1-seeds", type=int, default=100)
    ap.add_argument("--num-warmup", type=int, default=None)
    ap.add_argument("--num-samples", type=int, default=None)
    ap.add_argument("--thinning", type=int, default=None)
    ap.add_argument("--kernel", default="rwm", choices=["rwm", "sss"])
    a = ap.parse_args()
    out, _ = run(a.seeds, a.scale, a.w1_seeds, num_warmup=a.num_warmup, num_samples=a.num_samples, thinning=a.thinning, kernel=a.kernel)
    print(json.dumps(out))
