"""Time the GPU assignment (W1 of two 10^4-point samples) on draws shaped like the quality tables' inputs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np
import torch

from adaptive_mcmc_b200.utils import evaluation as ev

for n, d in ((10000, 10), (10000, 26), (4000, 10), (10001, 10)):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, d, device="cuda", generator=g)
    y = torch.randn(n, d, device="cuda", generator=g) * 1.05 + 0.02
    cm = ev.cost_matrix(x, y, 2.0)
    times = []
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r, c, info = ev.linear_sum_assignment(cm, return_info=True)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    dt = min(times[1:])
    print("   reps (ms):", " ".join(f"{t*1e3:.0f}" for t in times), flush=True)
    assert len(set(c.tolist())) == n
    print(f"n={n} d={d}: {dt*1e3:.1f} ms  rounds+bids={info['rounds']}  W1={info['cost_sum']/n:.6f}", flush=True)
    if n <= 4000:
        from scipy.optimize import linear_sum_assignment as lsa
        cmn = cm.cpu().numpy().astype(np.float64)
        rr, cc = lsa(cmn)
        print("   scipy W1", cmn[rr, cc].mean(), " gpu-on-float-costs", cmn[np.arange(n), c.cpu().numpy()].mean())
