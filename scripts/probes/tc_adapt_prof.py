import torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
data=models.synthetic_diamonds()
C=int(sys.argv[1]) if len(sys.argv)>1 else 65536
s=am.ARWMH(models.diamonds,num_chains=C); s.impl=_lib.IMPL_TENSOR
st=s.init(0,num_warmup=0,init_params=None,model_kwargs=data)
b=am.ChainBatch.from_state(s.potential,st,copy=False)
b.set_dense_scale(torch.eye(26)*0.002)
s.run_batch(b,30,collect=())
s.run_batch(b,30,collect=())
torch.cuda.synchronize(); print("ok", float(b.macc.mean()))
