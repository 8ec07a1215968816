"""Model families: the NumPyro model functions of the reference scripts, bound to data and
compiled into the CUDA kernels (the potential is inlined in the fused sampler step).

Each family mirrors one reference model (keyword names = the reference model's arguments,
so ``MCMC.run(key, **data)`` works unchanged):

  eight_schools(sigma, y)            python/scripts/run_eight_schools_lr_decay.py:26-35
  diamonds(Y, X)                     python/scripts/run_diamonds_lr_decay.py:24-40
  kidiq(mom_iq, mom_hs, kid_score)   python/scripts/run_kidiq_kidscore_lr_decay.py:29-41
  std_normal(d)                      python/jupyter/asumptions_check.ipynb cells 17-28
  gaussian(prec_chol)                BASELINE.json config 5 (not in the reference)

Flat layout of the unconstrained vector = ravel_pytree order of the site dict (sorted site
names), confirmed by python/scripts/eval_*.py.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib


def _np64(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _dtype_code(dtype):
    if dtype == torch.float32:
        return _lib.AMCMC_F32
    if dtype == torch.float64:
        return _lib.AMCMC_F64
    raise ValueError(f"dtype must be torch.float32 or torch.float64, got {dtype}")


class PotentialFn:
    """potential_fn(z) of the reference (arwmh.py:116,121,170): a model bound to its data, living
    on one CUDA device as an opaque amcmc_model handle."""

    def __init__(self, family, sites, arrays, dtype=torch.float32, device=None, data=None):
        self.family = family
        self.sites = OrderedDict(sites)  # name -> shape tuple (per chain)
        self.dim = int(sum(int(np.prod(s)) if len(s) else 1 for s in self.sites.values()))
        self.dtype = dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("adaptive_mcmc_b200 runs on CUDA devices only (no CPU fallback)")
        self.data = data or {}
        arrs = [_np64(a).ravel() for a in arrays]
        n = len(arrs)
        ptrs = (C.POINTER(C.c_double) * max(n, 1))()
        lens = (C.c_int64 * max(n, 1))()
        for k, a in enumerate(arrs):
            ptrs[k] = a.ctypes.data_as(C.POINTER(C.c_double))
            lens[k] = a.size
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = _lib.lib().amcmc_model_create(
                C.byref(self._handle), family.model_id, _dtype_code(dtype), self.dim, n, ptrs, lens
            )
        _lib.check(rc, f"amcmc_model_create({family.name})")
        self._keep = arrs

    @property
    def handle(self):
        return self._handle

    def __del__(self):
        try:  # may run during interpreter shutdown, when module globals are already gone
            h = getattr(self, "_handle", None)
            if h is not None and h.value:
                _lib.lib().amcmc_model_destroy(h)
                self._handle = None
        except Exception:
            pass

    # ---- flat <-> dict ----------------------------------------------------
    def ravel(self, z):
        """dict of [C, *shape] (or flat [C, d]) -> flat chain-major tensor [C, d]."""
        if isinstance(z, dict):
            parts = []
            for name, shape in self.sites.items():
                v = torch.as_tensor(z[name], dtype=self.dtype, device=self.device)
                n = int(np.prod(shape)) if len(shape) else 1
                if v.dim() == len(shape):  # no chain axis
                    v = v.unsqueeze(0)
                parts.append(v.reshape(v.shape[0], n))
            C_ = max(p.shape[0] for p in parts)
            parts = [p.expand(C_, p.shape[1]) for p in parts]
            return torch.cat(parts, dim=1).contiguous()
        v = torch.as_tensor(z, dtype=self.dtype, device=self.device)
        if v.dim() == 1:
            v = v.unsqueeze(0)
        if v.shape[-1] != self.dim:
            raise ValueError(f"expected last dimension {self.dim}, got {tuple(v.shape)}")
        return v.reshape(-1, self.dim).contiguous()

    def unravel(self, flat_cd):
        """flat [..., d] -> dict of [..., *shape] views."""
        out = OrderedDict()
        off = 0
        for name, shape in self.sites.items():
            n = int(np.prod(shape)) if len(shape) else 1
            v = flat_cd[..., off : off + n]
            out[name] = v.reshape(*flat_cd.shape[:-1], *shape) if len(shape) else v[..., 0]
            off += n
        return out

    def __call__(self, z):
        flat = self.ravel(z)  # [C, d]
        soa = flat.t().contiguous()  # [d, C]
        out = torch.empty(flat.shape[0], dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().amcmc_potential(
                self._handle, flat.shape[0], soa.data_ptr(), out.data_ptr(),
                C.c_void_p(torch.cuda.current_stream().cuda_stream),
            )
        _lib.check(rc, "amcmc_potential")
        return out

    def postprocess(self, z_dict):
        return self.family.postprocess(z_dict, self.data)


class ModelFamily:
    name = "?"
    model_id = -1

    def bind(self, *args, dtype=torch.float32, device=None, **kwargs):
        raise NotImplementedError

    def postprocess(self, z, data):
        return OrderedDict(z)

    def __repr__(self):
        return f"<model family {self.name}>"


class _EightSchools(ModelFamily):
    """def model(sigma, y=None) -- run_eight_schools_lr_decay.py:26.  Unconstrained sites
    (sorted): mu, tau (log), theta_base[8]."""

    name = "eight_schools"
    model_id = _lib.MODEL_EIGHT_SCHOOLS
    Y = (28, 8, -3, 7, -1, 1, 18, 12)
    SIGMA = (15, 10, 16, 11, 9, 11, 10, 18)

    def bind(self, sigma=None, y=None, dtype=torch.float32, device=None, **extra):
        sigma = self.SIGMA if sigma is None else sigma
        y = self.Y if y is None else y
        sites = [("mu", ()), ("tau", ()), ("theta_base", (8,))]
        return PotentialFn(self, sites, [_np64(y), _np64(sigma)], dtype, device, dict(sigma=sigma, y=y))

    def postprocess(self, z, data):
        # numpyro postprocess_fn: constrained sites + the deterministic `theta` of TransformReparam
        tau = torch.exp(z["tau"])
        out = OrderedDict()
        out["mu"] = z["mu"]
        out["tau"] = tau
        out["theta"] = z["mu"].unsqueeze(-1) + tau.unsqueeze(-1) * z["theta_base"]
        out["theta_base"] = z["theta_base"]
        return out


class _Kidiq(ModelFamily):
    """def model(mom_iq, mom_hs, kid_score=None) -- run_kidiq_kidscore_lr_decay.py:29.
    Sites (sorted): beta[3], sigma (log)."""

    name = "kidiq"
    model_id = _lib.MODEL_KIDIQ

    def bind(self, mom_iq, mom_hs, kid_score, dtype=torch.float32, device=None, **extra):
        sites = [("beta", (3,)), ("sigma", ())]
        return PotentialFn(self, sites, [_np64(kid_score), _np64(mom_hs), _np64(mom_iq)], dtype, device,
                           dict(mom_iq=mom_iq, mom_hs=mom_hs, kid_score=kid_score))

    def postprocess(self, z, data):
        return OrderedDict(beta=z["beta"], sigma=torch.exp(z["sigma"]))


class _Diamonds(ModelFamily):
    """def model(Y, X) -- run_diamonds_lr_decay.py:24.  Sites (sorted, 'I' < 'b'): Intercept, b[Kc], sigma (log)."""

    name = "diamonds"
    model_id = _lib.MODEL_DIAMONDS

    def bind(self, Y, X, dtype=torch.float32, device=None, **extra):
        X = _np64(X)
        Y = _np64(Y)
        if X.ndim != 2 or Y.shape != (X.shape[0],):
            raise ValueError("diamonds: X must be [N, K] and Y [N]")
        sites = [("Intercept", ()), ("b", (X.shape[1] - 1,)), ("sigma", ())]
        return PotentialFn(self, sites, [X, Y], dtype, device, dict(Y=Y, X=X))

    def postprocess(self, z, data):
        return OrderedDict(Intercept=z["Intercept"], b=z["b"], sigma=torch.exp(z["sigma"]))


class _StdNormal(ModelFamily):
    """potential_fn = 0.5*|x|^2 (asumptions_check.ipynb).  Site: x[d]."""

    name = "std_normal"
    model_id = _lib.MODEL_STD_NORMAL

    def bind(self, d=1, dtype=torch.float32, device=None, **extra):
        return PotentialFn(self, [("x", (int(d),))], [], dtype, device, dict(d=int(d)))


class _Gaussian(ModelFamily):
    """N(0, Sigma) with Sigma^-1 = P P^T (P lower Cholesky).  Site: x[d].  BASELINE.json config 5."""

    name = "gaussian"
    model_id = _lib.MODEL_GAUSSIAN

    def bind(self, prec_chol, dtype=torch.float32, device=None, **extra):
        P = _np64(prec_chol)
        if P.ndim != 2 or P.shape[0] != P.shape[1]:
            raise ValueError("gaussian: prec_chol must be square")
        return PotentialFn(self, [("x", (P.shape[0],))], [np.tril(P)], dtype, device, dict(prec_chol=P))


eight_schools = _EightSchools()
kidiq = _Kidiq()
diamonds = _Diamonds()
std_normal = _StdNormal()
gaussian = _Gaussian()

FAMILIES = {f.name: f for f in (eight_schools, kidiq, diamonds, std_normal, gaussian)}


# ---- synthetic data generators (posteriordb is not available offline) ----
def synthetic_diamonds(n=5000, k=25, seed=0, sigma=0.123):
    """Synthetic stand-in for posteriordb diamonds-diamonds (N=5000, K=25: column 0 ones, columns
    1-4 strongly collinear as in posteriordb_diamonds.ipynb cell 9, the rest centred contrasts),
    Y = X beta* + sigma*eps with Y in roughly [5.9, 9.8]."""
    rng = np.random.default_rng(seed)
    X = np.empty((n, k))
    X[:, 0] = 1.0
    base = rng.normal(size=n)
    X[:, 1] = base
    X[:, 2] = base**2 - 1.0 + 0.05 * rng.normal(size=n)
    X[:, 3] = base + 0.03 * rng.normal(size=n)
    X[:, 4] = base + 0.3 * rng.normal(size=n)
    for j in range(5, k):
        p = rng.uniform(0.1, 0.5)
        X[:, j] = (rng.random(n) < p).astype(np.float64)
    beta = np.zeros(k)
    beta[0] = 7.79
    beta[1:5] = [0.6, -0.05, 0.3, 0.1]
    beta[5:] = rng.normal(scale=0.15, size=k - 5)
    Y = X @ beta + sigma * rng.normal(size=n)
    return dict(X=X, Y=Y)


def diamonds_from_sufficient_stats(n, G, h, yy, seed=0, col_means=None):
    """A diamonds-shaped data set (X [n, K] with column 0 = ones, Y [n]) with the given sufficient statistics of the Gaussian
    linear likelihood: G = X1^T X1, h = X1^T Y, yy = Y^T Y for X1 = [1 | centred predictors] (python/scripts/run_diamonds_lr_decay.py:
    24-40 depends on the data through these only).  Xc = Q R with Q orthonormal and orthogonal to the ones vector, R^T R = G[1:,1:];
    Y = mean + Xc b_ols + r with r orthogonal to [1, Q] and |r|^2 = the residual sum of squares.  Used with the statistics recovered
    from the posteriordb reference draws the reference ships (tests/golden/reference_pins.json:diamonds_recovered_stats)."""
    G = np.asarray(G, np.float64)
    h = np.asarray(h, np.float64)
    k = G.shape[0]
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(np.column_stack([np.ones(n), rng.normal(size=(n, k))]))
    R = np.linalg.cholesky(G[1:, 1:]).T
    Xc = Q[:, 1:k] @ R
    beta = np.linalg.solve(G, h)
    rss0 = float(yy) - h @ beta
    Y = beta[0] + Xc @ beta[1:] + np.sqrt(max(rss0, 0.0)) * Q[:, k]
    if col_means is None:
        col_means = np.linspace(0.2, 1.5, k - 1)
    return dict(X=np.column_stack([np.ones(n), Xc + np.asarray(col_means)[None, :]]), Y=Y)


def synthetic_kidiq(n=434, seed=0):
    rng = np.random.default_rng(seed)
    mom_hs = (rng.random(n) < 0.79).astype(np.float64)
    mom_iq = rng.normal(100.0, 15.0, size=n)
    kid = 26.0 + 6.0 * mom_hs + 0.55 * mom_iq + 18.0 * rng.normal(size=n)
    return dict(mom_iq=mom_iq, mom_hs=mom_hs, kid_score=kid)


def ar1_precision_chol(d=200, rho=0.9):
    """Lower Cholesky factor P of the precision of Sigma_ij = rho^|i-j|."""
    # the AR(1) precision is tridiagonal: Q = (1-rho^2)^-1 * tridiag(-rho, [1, 1+rho^2, ..., 1+rho^2, 1], -rho);
    # its Cholesky factor is exactly lower-bidiagonal
    Q = np.zeros((d, d))
    i = np.arange(d)
    Q[i, i] = 1.0 + rho * rho
    Q[0, 0] = Q[-1, -1] = 1.0
    Q[i[1:], i[:-1]] = Q[i[:-1], i[1:]] = -rho
    Q /= 1.0 - rho * rho
    P = np.linalg.cholesky(Q)
    P[np.abs(P) < 1e-14] = 0.0
    return P
