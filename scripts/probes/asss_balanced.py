"""ASSS eight_schools, 65,536 chains x 2,000 steps: plain thread-per-chain kernel (impl 1) against the work-queue launch (impl 4)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import _lib, models

Cn, T = 65536, 2000
for impl in (_lib.IMPL_REGISTER, _lib.IMPL_REGISTER_BALANCED):
    s = am.ASSS(models.eight_schools, num_chains=Cn)
    s.impl = impl
    b = s._batch_from_state(s.init(0, num_warmup=0, init_params=None))
    s.run_batch(b, T, collect=())
    torch.cuda.synchronize()
    ms = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.run_batch(b, T, thinning=50); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    print(f"impl {impl}: {ms[len(ms) // 2]:.2f} ms  {Cn * T / ms[len(ms) // 2] / 1e-3:.3e} chain-steps/s  mean shrink iterations {float(b.macc.mean()):.3f}")
