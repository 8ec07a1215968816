import numpy as np, torch, sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models
data=models.synthetic_diamonds()
for C in (4096, 65536):
    s=am.ARWMH(models.diamonds,num_chains=C)
    st=s.init(0,num_warmup=0,init_params=None,model_kwargs=data)
    b=am.ChainBatch.from_state(s.potential,st,copy=False)
    s.run_batch(b,5,collect=())
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    T=20
    e0.record(); s.run_batch(b,T,collect=()); e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1); print(C,'chains block kernel', ms,'ms','%.3g chain-steps/s'%(C*T/ms*1e3))
