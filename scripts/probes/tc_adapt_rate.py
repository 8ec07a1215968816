"""Rate of the per-chain-adaptive diamonds tensor-core kernel at the bench launch shape (65,536 chains x 500 steps).
Environment hooks of run_diamonds_tc_adapt (AMCMC_TC_*) apply; prints one line."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models

C = int(os.environ.get("CHAINS", 65536)); T = 500
data = models.synthetic_diamonds(n=5000, k=25, seed=0)
X, Y = data["X"], data["Y"]
Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
q0 = mode[None] + 0.004 * np.random.default_rng(0).normal(size=(C, 26))
s = am.ARWMH(models.diamonds, num_chains=C, init_strategy=am.init_to_value(torch.from_numpy(q0)))
st = s.init(0, num_warmup=0, init_params=None, model_kwargs=data)
b = am.ChainBatch.from_state(s.potential, st, copy=False)
b.set_dense_scale(torch.eye(26) * 0.002)
for _ in range(2):
    s.run_batch(b, T, collect=())
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(int(os.environ.get('REPS', 4)))]
for e0, e1 in ev:
    e0.record(); s.run_batch(b, T, collect=()); e1.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(c) for a, c in ev)
print(json.dumps({"tag": os.environ.get("TAG", ""), "chains": C, "ms_median": ms[len(ms) // 2], "rate": C * T / (ms[len(ms) // 2] * 1e-3),
                  "frac_3pass": C * T / (ms[len(ms) // 2] * 1e-3) * 720000 / 1377.6e12, "macc": float(b.macc.mean()),
                  "env": {k: v for k, v in os.environ.items() if k.startswith("AMCMC_TC")}}))
