"""Time the three kernel sums of the MMD estimators: one tcgen05 launch against the three CUDA-core passes."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from adaptive_mcmc_b200.utils import evaluation as ev

rng = np.random.default_rng(0)
for n, d in ((10000, 26), (10000, 10), (10000, 4), (40000, 26), (2048, 26)):
    x = torch.from_numpy(rng.normal(size=(n, d)).astype(np.float32)).cuda()
    y = torch.from_numpy((rng.normal(size=(n, d)) + 0.1).astype(np.float32)).cuda()
    g = 4.0 / ev.sqdist_median(y[:4000])
    for impl in ("tc", "cuda"):
        ev.mmd_kernel_sums(x, y, g, impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(10):
            r = ev.mmd_kernel_sums(x, y, g, impl)
        e1.record()
        torch.cuda.synchronize()
        print(f"n=m={n} d={d} {impl:5s}: {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us per call (wall {(time.perf_counter() - t0) / 10 * 1e6:8.1f} us)  sums {r}")
    t0 = time.perf_counter()
    for _ in range(5):
        v = ev.mmd_heuristic(x, y)
    torch.cuda.synchronize()
    print(f"   mmd_heuristic {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms  -> {v:.6f}")
