import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import adaptive_mcmc_b200 as am
from eval_eight_schools import unconstrained
from adaptive_mcmc_b200 import diagnostics
np.set_printoptions(precision=3, suppress=True, linewidth=200)
out = {}
for name, smp, cfg in (("rwm", am.ARWMH(am.models.eight_schools), (50000, 500000, 50)), ("sss", am.ASSS(am.models.eight_schools), (25000, 250000, 25)),
                       ("sss64", am.ASSS(am.models.eight_schools, dtype=torch.float64), (25000, 250000, 25))):
    mcmc = am.MCMC(smp, num_warmup=cfg[0], num_samples=cfg[1], thinning=cfg[2], num_chains=100)
    mcmc.run(0)
    x = unconstrained(mcmc.get_samples(group_by_chain=True)).double()
    ess = torch.stack([diagnostics.effective_sample_size(x[k:k+1]) for k in range(0, 100, 10)])
    print(name, "pooled mean", x.reshape(-1, 10).mean(0).cpu().numpy())
    print(name, "pooled sd  ", x.reshape(-1, 10).std(0).cpu().numpy())
    print(name, "ESS per chain (10 chains): min %.0f median %.0f max %.0f" % (ess.min(), ess.median(), ess.max()))
    print(name, "per-seed mean sd", x.mean(1).std(0).cpu().numpy())
