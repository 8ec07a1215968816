#!/usr/bin/env python
"""Runs the reference-style configurations of BASELINE.json (configs[0] and configs[1]) through the
drop-in interface, the way the reference's own scripts do (python/scripts/run_eight_schools_wasserstein.py:37-70,
run_diamonds_wasserstein.py:42-62), and prints wall times + posterior summaries.

  configs[0]: eight_schools (d=10), adaptive Metropolis, 4 chains x 10k steps   (+1k warm-up)
  configs[1]: diamonds (d=26, N=5000), adaptive Metropolis, 64 chains x 50k steps
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_mcmc_b200 as am  # noqa: E402
from adaptive_mcmc_b200 import models  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def config0(dtype):
    sampler = am.ARWMH(models.eight_schools, dtype=dtype)
    mcmc = am.MCMC(sampler, num_warmup=1000, num_samples=10000, num_chains=4)
    _, el = timed(lambda: mcmc.run(0, sigma=models.eight_schools.SIGMA, y=models.eight_schools.Y,
                                   extra_fields=("potential_energy",)))
    s = mcmc.get_samples(group_by_chain=True)
    return dict(config="eight_schools 4 chains x (1k + 10k) steps", dtype=str(dtype), wall_s=el,
                chain_steps_per_s=4 * 11000 / el, mu_mean=float(s["mu"].mean()), tau_mean=float(s["tau"].mean()),
                accept=float(mcmc.last_state.mean_accept_prob.mean()))


def config1(dtype):
    data = models.synthetic_diamonds(n=5000, k=25, seed=0)
    sampler = am.ARWMH(models.diamonds, dtype=dtype)
    mcmc = am.MCMC(sampler, num_warmup=0, num_samples=50000, thinning=50, num_chains=64)
    _, el = timed(lambda: mcmc.run(0, **data, extra_fields=("potential_energy",)))
    pe = mcmc.get_extra_fields(group_by_chain=True)["potential_energy"]
    return dict(config="diamonds (synthetic, N=5000, d=26) 64 chains x 50k steps", dtype=str(dtype), wall_s=el,
                chain_steps_per_s=64 * 50000 / el, first_U=float(pe[:, 0].mean()), last_U=float(pe[:, -1].mean()),
                accept=float(mcmc.last_state.mean_accept_prob.mean()))


if __name__ == "__main__":
    assert torch.cuda.is_available()
    out = []
    for dt in (torch.float32, torch.float64):
        config0(dt)  # warm-up (module load, allocator)
        out.append(config0(dt))
        out.append(config1(dt))
    for r in out:
        print(json.dumps(r))
