"""Two launches of the per-chain-adaptive diamonds tensor-core kernel at the bench launch shape (65,536 chains, 256-step
segments) for ncu: `ncu --set full -k regex:diamonds_tc_adapt_kernel -s 2 -c 1 python scripts/probes/tc_adapt_prof.py`."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models
C, T = 65536, 256
data = models.synthetic_diamonds(n=5000, k=25, seed=0)
X, Y = data["X"], data["Y"]
Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
q0 = mode[None] + 0.004 * np.random.default_rng(0).normal(size=(C, 26))
s = am.ARWMH(models.diamonds, num_chains=C, init_strategy=am.init_to_value(torch.from_numpy(q0)))
st = s.init(0, num_warmup=0, init_params=None, model_kwargs=data)
b = am.ChainBatch.from_state(s.potential, st, copy=False)
b.set_dense_scale(torch.eye(26) * 0.002)
for _ in range(4):
    s.run_batch(b, T, collect=())
torch.cuda.synchronize()
print("ok", float(b.macc.mean()))
