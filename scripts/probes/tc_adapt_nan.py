import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
data = models.synthetic_diamonds()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
s = am.ARWMH(models.diamonds, num_chains=C); s.impl = _lib.IMPL_TENSOR
st = s.init(0, num_warmup=0, init_params=None, model_kwargs=data)
b = am.ChainBatch.from_state(s.potential, st, copy=False)
b.set_dense_scale(torch.eye(26) * 0.002)
print("init pe finite:", bool(torch.isfinite(b.pe).all()), "max pe %.3g" % float(b.pe.max()))
for k in range(3):
    s.run_batch(b, 50, collect=())
    bad = ~torch.isfinite(b.macc) | ~torch.isfinite(b.pe) | ~torch.isfinite(b.z).all(0)
    idx = torch.nonzero(bad).flatten().cpu().numpy()
    print("after", (k + 1) * 50, "steps: bad chains", len(idx), "groups", np.unique(idx // 128)[:20], "cta (148)", np.unique((idx // 128) % 148)[:20], "local", np.unique((idx // 128) // 148))
    if len(idx):
        c = idx[0]; print("  example chain", c, "pe", float(b.pe[c]), "macc", float(b.macc[c]), "lam", float(b.lam[c]), "z", b.z[:, c].cpu().numpy()[[0, 1, 25]])
