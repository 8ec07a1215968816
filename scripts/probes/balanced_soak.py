"""Long launches through the per-SM work queue: 65,536 eight_schools chains x 10^6 steps in ONE launch (1,953 hand-offs per group),
then the same run cut in two launches; prints rates and the posterior means (mu ~ 4.4, tau ~ 3.6 for the centred model)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models

Cn = 65536
s = am.ARWMH(models.eight_schools, num_chains=Cn)
b = s._batch_from_state(s.init(1, num_warmup=100_000, init_params=None))
for T in (1_000_000, 500_000, 500_000):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    raw = s.run_batch(b, T, thinning=T // 4)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    z = raw["z"]                      # [S, d, C]
    print(f"T={T}: {dt:.3f} s  {Cn * T / dt:.3e} chain-steps/s  samples {tuple(z.shape)}  mu {float(z[-1, 0].mean()):.3f}  "
          f"log tau {float(z[-1, 1].mean()):.3f}  accept {float(b.macc.mean()):.4f}  finite {bool(torch.isfinite(z).all())}", flush=True)
