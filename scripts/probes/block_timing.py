import sys, os, ctypes; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
L = _lib.lib()
data = models.synthetic_diamonds()
names = ["loop top/collect", "draws+sync", "proposal matvec", "potential (likelihood)", "accept+mean+sync_and", "sweep+sync"]
for C in (64, 148, 200, 296, 400):
    s = am.ARWMH(models.diamonds, num_chains=C); s.impl = _lib.IMPL_BLOCK
    st = s.init(0, num_warmup=0, init_params=None, model_kwargs=data)
    b = am.ChainBatch.from_state(s.potential, st, copy=False)
    b.set_dense_scale(torch.eye(26) * 0.002)
    s.run_batch(b, 2000, collect=()); torch.cuda.synchronize()
    T = 5000
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); s.run_batch(b, T, collect=()); e1.record(); torch.cuda.synchronize()
    print(C, "chains: %.2f us/step" % (e0.elapsed_time(e1) / T * 1e3))
