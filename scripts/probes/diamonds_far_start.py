import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
from oracle import arwmh_numpy as o
data = models.synthetic_diamonds()
pot = o.make_potential("diamonds", **data)
C = 4096
for impl, name in ((_lib.IMPL_TENSOR, "tensor"), (_lib.IMPL_BLOCK, "block")):
    s = am.ARWMH(models.diamonds, num_chains=C); s.impl = impl
    st = s.init(11, num_warmup=0, init_params=None, model_kwargs=data)
    b = am.ChainBatch.from_state(s.potential, st)
    T = 20000 if impl == _lib.IMPL_TENSOR else 4000
    raw = s.run_batch(b, T, thinning=1000, collect=("z", "potential_energy"))
    sel = np.arange(0, C, 16)
    z = raw["z"].double().cpu().numpy()[:, :, sel]; pe = raw["potential_energy"].double().cpu().numpy()[:, sel]
    exact = np.stack([pot(z[k].T) for k in range(z.shape[0])])
    for k in range(z.shape[0]):
        print(name, "step", (k + 1) * 1000, "median U %.4g  min U %.4g  max|err| %.3g  max rel %.2g  macc %.3f lam %.2f" % (
            np.median(exact[k]), exact[k].min(), np.abs(pe[k] - exact[k]).max(), (np.abs(pe[k] - exact[k]) / np.maximum(1, np.abs(exact[k]))).max(),
            float(b.macc.mean()), float(b.lam.mean())))
