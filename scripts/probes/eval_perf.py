import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from adaptive_mcmc_b200.utils import evaluation as ev
rng = np.random.default_rng(0)
n, d = 10000, 26
x = torch.from_numpy(rng.normal(size=(n, d)).astype(np.float32)).cuda()
y = torch.from_numpy((rng.normal(size=(n, d)) + 0.1).astype(np.float32)).cuda()
for name, fn in (("sqdist_median", lambda: ev.sqdist_median(y)), ("kernel_sum", lambda: ev.gaussian_kernel_sum(x, y, 0.02)),
                 ("mmd_heuristic", lambda: ev.mmd_heuristic(x, y)), ("cost_matrix", lambda: ev.cost_matrix(x, y)),
                 ("moment_rmse", lambda: ev.pth_moment_rmse(x, y)), ("max_sliced_w (1000 dirs)", lambda: ev.max_sliced_wasserstein(x, y, 0))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): r = fn()
    torch.cuda.synchronize()
    print("%-26s %8.3f ms" % (name, (time.perf_counter() - t0) / 5 * 1e3), r if not isinstance(r, torch.Tensor) else r.shape)
