import numpy as np, torch, sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
data=models.synthetic_diamonds()
for C in (65536, 131072):
    for impl,name,T in ((_lib.IMPL_TENSOR,'tc-adapt',400),(_lib.IMPL_BLOCK,'block',20)):
        s=am.ARWMH(models.diamonds,num_chains=C); s.impl=impl
        st=s.init(0,num_warmup=0,init_params=None,model_kwargs=data)
        b=am.ChainBatch.from_state(s.potential,st,copy=False)
        b.set_dense_scale(torch.eye(26)*0.002)
        s.run_batch(b,T,collect=())
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); s.run_batch(b,T,collect=()); e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1); print(C,name, '%.2f ms'%ms,'%.3g chain-steps/s'%(C*T/ms*1e3), 'macc %.3f'%float(b.macc.mean()), flush=True)
