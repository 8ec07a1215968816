import torch, sys, ctypes, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
L=_lib.lib()
data=models.synthetic_diamonds()
names=["top(other)","epilogue(GEMM)","exchange","accept+loc","v_full wait","pass","propose+emit"]
for C in (16384, 65536):
    s=am.ARWMH(models.diamonds,num_chains=C); s.impl=_lib.IMPL_TENSOR
    st=s.init(0,num_warmup=0,init_params=None,model_kwargs=data)
    b=am.ChainBatch.from_state(s.potential,st,copy=False)
    b.set_dense_scale(torch.eye(26)*0.002)
    s.run_batch(b,100,collect=()); torch.cuda.synchronize()
    buf=(ctypes.c_ulonglong*16)()
    L.amcmc_debug_tc_timing(buf,1)
    T=200
    s.run_batch(b,T,collect=()); torch.cuda.synchronize()
    L.amcmc_debug_tc_timing(buf,0)
    v=np.array(list(buf)[:7],dtype=np.float64)
    print(C,"cycles/step by phase (thread 0 of block 0; both owned chains):")
    for n,x in zip(names,v): print("   %-16s %9.0f  %5.1f%%"%(n,x/T,100*x/v.sum()))
    print("   total %.0f cycles/step"%(v.sum()/T))
