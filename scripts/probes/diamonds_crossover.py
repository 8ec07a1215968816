import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
data = models.synthetic_diamonds()
for C in (128, 256, 384, 512, 1024, 2048, 4096):
    out = []
    for impl, name in ((_lib.IMPL_TENSOR, "tc"), (_lib.IMPL_BLOCK, "block")):
        s = am.ARWMH(models.diamonds, num_chains=C); s.impl = impl
        st = s.init(0, num_warmup=0, init_params=None, model_kwargs=data)
        b = am.ChainBatch.from_state(s.potential, st, copy=False)
        b.set_dense_scale(torch.eye(26) * 0.002)
        T = 300
        s.run_batch(b, T, collect=())
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.run_batch(b, T, collect=()); e1.record(); torch.cuda.synchronize()
        out.append("%s %.1f us/step" % (name, e0.elapsed_time(e1) / T * 1e3))
    print(C, "chains:", " | ".join(out), flush=True)
