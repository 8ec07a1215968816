import numpy as np, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models
d,C=200,4096
P=models.ar1_precision_chol(d,0.9)
s=am.RAM(models.gaussian,num_chains=C,init_strategy=am.init_to_value(torch.zeros(C,d)))
st=s.init(3,num_warmup=0,init_params=None,model_kwargs=dict(prec_chol=P))
b=am.ChainBatch.from_state(s.potential,st,copy=False)
for T in (50,50):
    s.run_batch(b,T,collect=())
torch.cuda.synchronize()
