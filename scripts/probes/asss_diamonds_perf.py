import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models
data = models.synthetic_diamonds()
X, Y = data["X"], data["Y"]
Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
for C in (1, 64, 1024, 4096):
    q0 = mode[None] + 0.01 * np.random.default_rng(0).normal(size=(C, 26))
    s = am.ASSS(models.diamonds, num_chains=C, init_strategy=am.init_to_value(torch.from_numpy(q0)))
    st = s.init(0, num_warmup=0, init_params=None, model_kwargs=data)
    b = s._batch_from_state(st); b.set_dense_scale(torch.eye(26) * 0.01)
    T = 2000 if C <= 64 else 300
    s.run_batch(b, T, collect=())
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); s.run_batch(b, T, collect=()); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(C, "chains: %.1f ms for %d steps -> %.3g chain-steps/s (%.0f it/s per chain), mean shrink iterations %.2f" % (ms, T, C * T / ms * 1e3, T / ms * 1e3, float(b.macc.mean())))
