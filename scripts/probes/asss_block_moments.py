import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models
np.set_printoptions(precision=3, suppress=True, linewidth=200)
d = 12
P = models.ar1_precision_chol(d, 0.6)
Pn = np.asarray(P, np.float64)
cov = np.linalg.inv(np.tril(Pn) @ np.tril(Pn).T)
for dt in (torch.float64, torch.float32):
    for nw, T in ((2000, 6000), (20000, 60000)):
        s2 = am.ASSS(models.gaussian, num_chains=512, dtype=dt)
        st2 = s2.init(9, num_warmup=nw, init_params=None, model_kwargs=dict(prec_chol=P))
        coll2, last = s2.run(st2, T, thinning=20, collect_start=nw)
        x = coll2["z"]["x"].double().reshape(-1, d).cpu().numpy()
        emp = np.cov(x.T)
        print(dt, nw, T, "diag emp", np.diag(emp)[:6], "max err", np.abs(emp - cov).max(), "mean iters", float(last.mean_accept_prob.mean()) if hasattr(last, "mean_accept_prob") else None)
        b = s2._batch_from_state(last)
        print("   shrink iters mean", float(b.macc.mean()), " scale diag chain0", torch.diagonal(last.adapt_state.scale[0]).cpu().numpy()[:6])
# same target through ARWMH on the block kernel
s3 = am.ARWMH(models.gaussian, num_chains=512, dtype=torch.float32)
st3 = s3.init(9, num_warmup=20000, init_params=None, model_kwargs=dict(prec_chol=P))
coll3, _ = s3.run(st3, 60000, thinning=20, collect_start=20000)
x = coll3["z"]["x"].double().reshape(-1, d).cpu().numpy()
print("ARWMH diag emp", np.diag(np.cov(x.T))[:6], "max err", np.abs(np.cov(x.T) - cov).max())
