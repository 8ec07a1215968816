"""Few-chain diamonds (BASELINE configs[1] shape: 64 chains): one 512-thread CTA per chain against a 2-CTA cluster per chain.
AMCMC_BLOCK_CLUSTER=0/2/4/8 selects the cluster size."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import _lib, models

data = models.synthetic_diamonds(n=5000, k=25, seed=0)
for dt in (torch.float32, torch.float64):
    for C in (1, 18, 37, 64):
        s = am.ARWMH(models.diamonds, num_chains=C, dtype=dt)
        s.impl = _lib.IMPL_BLOCK
        b = s._batch_from_state(s.init(0, num_warmup=0, init_params=None, model_kwargs=data))
        s.run_batch(b, 200, collect=())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        raw = s.run_batch(b, 5000, thinning=100)
        torch.cuda.synchronize()
        dtm = time.perf_counter() - t0
        print(f"{str(dt)[6:]} chains {C}: {dtm * 1e6 / 5000:.2f} us per step  U {float(b.pe.mean()):.3f}  accept {float(b.macc.mean()):.4f}  "
              f"cluster={os.environ.get('AMCMC_BLOCK_CLUSTER', 'auto')}", flush=True)
