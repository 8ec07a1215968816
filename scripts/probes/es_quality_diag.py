import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import adaptive_mcmc_b200 as am
from eval_eight_schools import reference_draws, unconstrained
from adaptive_mcmc_b200.utils import evaluation as ev
np.set_printoptions(precision=3, suppress=True, linewidth=200)
for steps in (30000, 200000):
    y = reference_draws(steps=steps)
    print("y steps", steps, "mean", y.mean(0).cpu().numpy(), "\n   sd", y.std(0).cpu().numpy())
mcmc = am.MCMC(am.ARWMH(am.models.eight_schools), num_warmup=50000, num_samples=500000, thinning=50, num_chains=100)
mcmc.run(0)
x = unconstrained(mcmc.get_samples(group_by_chain=True))
print("x pooled mean", x.reshape(-1, 10).mean(0).cpu().numpy(), "\n   sd", x.reshape(-1, 10).std(0).cpu().numpy())
xm = x.mean(1)
print("per-seed mean sd", xm.std(0).cpu().numpy())
# independent check: two halves of seeds against each other
a, b = x[0].contiguous(), x[1].contiguous()
print("x0 vs x1: rmse", ev.pth_moment_rmse(a, b, 1), "mmd", ev.mmd_heuristic(a, b))
print("x0 vs y: rmse", ev.pth_moment_rmse(a, y, 1), "mmd", ev.mmd_heuristic(a, y))
yy = x[:, ::100, :].reshape(-1, 10).contiguous()   # 100 seeds x 100 widely spaced draws = 10^4 nearly independent draws
print("x0 vs pooled-thinned x: rmse", ev.pth_moment_rmse(a, yy, 1), "mmd", ev.mmd_heuristic(a, yy))
print("x5 vs pooled-thinned x: rmse", ev.pth_moment_rmse(x[5].contiguous(), yy, 1), "mmd", ev.mmd_heuristic(x[5].contiguous(), yy))
