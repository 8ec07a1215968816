"""Latency of the reference-style few-chain runs (BASELINE configs[0]: eight_schools, 4 chains x (1k + 10k) steps)."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
for C in (1, 4, 100, 1024):
    ts = []
    for k in range(6):
        mcmc = am.MCMC(am.ARWMH(am.models.eight_schools), num_warmup=1000, num_samples=10000, num_chains=C)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mcmc.run(k); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(json.dumps({"chains": C, "ms_median": 1e3 * sorted(ts)[3], "us_per_step": 1e6 * sorted(ts)[3] / 11000}))
