"""ASSS eight_schools, few chains: one warp per 32 chains against the producer/consumer pair (AMCMC_SMALL_DUO=0/1)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models

for C in (1, 100, 4736):
    s = am.ASSS(models.eight_schools, num_chains=C)
    b = s._batch_from_state(s.init(0, num_warmup=0, init_params=None))
    s.run_batch(b, 2000, collect=())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run_batch(b, 20000, thinning=100)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"chains {C}: {dt * 1e6 / 20000:.3f} us per step  duo={os.environ.get('AMCMC_SMALL_DUO', 'auto')}", flush=True)
