import torch, sys, ctypes, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
L=_lib.lib()
data=models.synthetic_diamonds()
names=["top(other)","epilogue(GEMM)","exchange","accept+loc","v_full wait","pass","propose+emit"]
for C in (65536,):
    s=am.ARWMH(models.diamonds,num_chains=C); s.impl=_lib.IMPL_TENSOR
    st=s.init(0,num_warmup=0,init_params=None,model_kwargs=data)
    b=am.ChainBatch.from_state(s.potential,st,copy=False)
    b.set_dense_scale(torch.eye(26)*0.002)
    s.run_batch(b,100,collect=()); torch.cuda.synchronize()
    buf=(ctypes.c_ulonglong*64)()
    L.amcmc_debug_tc_timing(buf,1)
    T=200
    s.run_batch(b,T,collect=()); torch.cuda.synchronize()
    L.amcmc_debug_tc_timing(buf,0)
    w=np.array(list(buf)[16:48],dtype=np.float64).reshape(8,4)/T
    print(C,"cycles/step per sampler warp (lane 0): GEMM phase | accept | v_full+pass | proposal+top | total")
    for k in range(8): print("   warp %d (stream %d): %8.0f %8.0f %8.0f %8.0f   %8.0f"%(k,k//4,*w[k],w[k].sum()))
    dr=np.array(list(buf)[48:64],dtype=np.float64).reshape(8,2)/T
    print("   drain loop per warp, cycles/step: waiting for acc_full | draining (fence, ld, FMA, fence, arrive):", " ".join("%.0f|%.0f"%tuple(x) for x in dr))
    m=np.array(list(buf)[8:12],dtype=np.float64)
    print("   MMA thread of block 0, cycles/step: x_full wait %.0f, a_ready wait %.0f, acc_empty wait %.0f, issue/other %.0f, total %.0f"%(m[0]/T,m[1]/T,m[2]/T,m[3]/T,m.sum()/T))
