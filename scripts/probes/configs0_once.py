"""BASELINE configs[0] once (eight_schools, 4 chains, 1,000 warm-up + 10,000 sampling steps through MCMC.run): the command of
profiles/r02_configs0_launches.csv."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am

for k in range(3):
    mcmc = am.MCMC(am.ARWMH(am.models.eight_schools), num_warmup=1000, num_samples=10000, num_chains=4)
    mcmc.run(k)
torch.cuda.synchronize()
print("ok")
