import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
from oracle import arwmh_numpy as o
np.set_printoptions(precision=4, suppress=True, linewidth=220)
data = models.synthetic_diamonds()
pot = o.make_potential("diamonds", **data)
off, j = 83328, 11
res = {}
for impl, name in ((_lib.IMPL_TENSOR, "tc"), (_lib.IMPL_BLOCK, "blk")):
    s = am.ARWMH(models.diamonds, num_chains=128, chain_offset=off); s.impl = impl
    st = s.init(0, num_warmup=0, init_params=None, model_kwargs=data)
    b = am.ChainBatch.from_state(s.potential, st, copy=False)
    b.set_dense_scale(torch.eye(26) * 0.002)
    raw = s.run_batch(b, 50, collect=("z", "potential_energy"), record_accept=True)
    res[name] = (raw["z"][:, :, j].double().cpu().numpy(), raw["potential_energy"][:, j].double().cpu().numpy(), raw["accept"][:, j].cpu().numpy())
zt, pt, at = res["tc"]; zb, pb, ab = res["blk"]
for t in range(50):
    ex = pot(zt[t][None])[0]
    print(t, "acc tc/blk", int(at[t]), int(ab[t]), "pe tc %.6g exact(at tc pos) %.6g | blk %.6g" % (pt[t], ex, pb[t]), " s tc %.4f blk %.4f  |z-zb| %.3g" % (zt[t, 25], zb[t, 25], np.abs(zt[t] - zb[t]).max()))
    if not np.isfinite(pt[t]): break
