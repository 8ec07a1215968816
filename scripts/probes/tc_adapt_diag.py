import sys, numpy as np, torch
sys.path.insert(0, '.')
import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200 import models, _lib
from oracle import arwmh_numpy as o, c_oracle as co
co.build()
data = models.synthetic_diamonds(seed=0) if hasattr(models, 'synthetic_diamonds') else None
X, Y = data["X"], data["Y"]
def mode(dd):
    X, Y = np.asarray(dd["X"], np.float64), np.asarray(dd["Y"], np.float64)
    Xc = X[:, 1:] - X[:, 1:].mean(0)
    b = np.linalg.solve(Xc.T @ Xc + 0.015 * np.eye(24), Xc.T @ (Y - Y.mean()))
    r = Y - Y.mean() - Xc @ b
    return np.concatenate([[Y.mean()], b, [np.log(r.std())]])
import importlib.util
spec = importlib.util.spec_from_file_location("tt", "tests/test_gpu_tc.py")
C, d, Tmax, nw = 1000, 26, 40, 10
rng = np.random.default_rng(3)
try:
    sys.path.insert(0, 'tests'); import test_gpu_tc as tt; q_mode = tt._mode(data)
except Exception as e:
    print("fallback mode", e); q_mode = mode(data)
q0 = q_mode[None] + 0.004 * rng.normal(size=(C, d))
nrm = rng.normal(size=(Tmax, C, d)).astype(np.float32)
uni = rng.random(size=(Tmax, C)).astype(np.float32)
pot = o.make_potential("diamonds", **data)
rows = []
for T in [1, 2, 3, 5, 8, 10, 11, 12, 15, 20, 30, 40]:
    s = am.ARWMH(models.diamonds, num_chains=C, init_strategy=am.init_to_value(torch.from_numpy(q0)))
    s.impl = _lib.IMPL_TENSOR
    st = s.init(1, num_warmup=nw, init_params=None, model_kwargs=data)
    b = am.ChainBatch.from_state(s.potential, st); b.set_dense_scale(torch.eye(d) * 0.002); st = b.to_state()
    z0 = b.z.t().double().cpu().numpy()
    ost = o.ARWMHState(0, z0, pot(z0), np.zeros(C), o.ARWMHAdaptState(z0.copy(), np.broadcast_to(np.eye(d) * 0.002, (C, d, d)).copy(), np.zeros(C)), np.zeros(C), 0)
    coll, last = s.run(st, T, draws=(torch.from_numpy(nrm[:T]), torch.from_numpy(uni[:T])), record_accept=True)
    olast, ocoll = co.arwmh_run(ost, "diamonds", T, draws=(nrm[:T].astype(np.float64), uni[:T].astype(np.float64)), record_accept=True, num_warmup=nw, **data)
    same = (coll["accept"].cpu().numpy() == ocoll["accepts"]).all(axis=0)
    dl = np.abs(last.adapt_state.log_step_size.cpu().numpy() - olast.adapt_state.log_step_size)
    dpe = np.abs(last.potential_energy.cpu().numpy() - olast.potential_energy)
    dm = np.abs(last.mean_accept_prob.cpu().numpy() - olast.mean_accept_prob)
    w = np.argmax(np.where(same, dl, 0))
    print(f"T={T:3d} same={same.mean():.3f} dlam med={np.median(dl[same]):.2e} max={dl[same].max():.2e} (chain {w}) dpe med={np.median(dpe[same]):.2e} max={dpe[same].max():.2e} dmacc max={dm[same].max():.2e}"
          f" | worst: lam={last.adapt_state.log_step_size[w].item():.4f}/{olast.adapt_state.log_step_size[w]:.4f} pe={last.potential_energy[w].item():.4f}/{olast.potential_energy[w]:.4f} acc={coll['accept'][:, w].int().tolist()[-6:]}")
if True:
    pe0 = b.pe.double().cpu().numpy() if hasattr(b, 'pe') else None
    print("init pe err", np.abs(pe0 - pot(z0)).max())
