"""diamonds ASSS, few chains: CTA per chain against a cluster per chain (AMCMC_BLOCK_CLUSTER=0 forces the former)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models

data = models.synthetic_diamonds(n=5000, k=25, seed=0)
for C in (1, 18, 64):
    s = am.ASSS(models.diamonds, num_chains=C)
    b = s._batch_from_state(s.init(0, num_warmup=0, init_params=None, model_kwargs=data))
    s.run_batch(b, 200, collect=())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run_batch(b, 3000, thinning=100)
    torch.cuda.synchronize()
    dtm = time.perf_counter() - t0
    print(f"chains {C}: {dtm * 1e6 / 3000:.2f} us per step ({3000 / dtm:.0f} it/s per chain)  U {float(b.pe.mean()):.3f}  mean shrink iterations "
          f"{float(b.macc.mean()):.3f}  cluster={os.environ.get('AMCMC_BLOCK_CLUSTER', 'auto')}", flush=True)
