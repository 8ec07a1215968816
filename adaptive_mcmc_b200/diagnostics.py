"""Chain diagnostics consumed by the reference's notebooks through ``mcmc.print_summary()``:
n_eff (Geyer initial-monotone-sequence ESS) and split R-hat, restated from
``numpyro.diagnostics`` (3rd-party; same definitions so min-ESS/s is comparable with the
reference's recorded table, posteriordb_eight-schools.ipynb:L855-865).  Runs on whatever
device the samples live on (torch.fft); this is the caller side of the hot path, not the hot path.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch


def _next_fast_len(target):
    if target <= 2:
        return target
    while True:
        m = target
        for p in (2, 3, 5):
            while m % p == 0:
                m //= p
        if m == 1:
            return target
        target += 1


def autocorrelation(x, axis=0):
    x = torch.as_tensor(x).double().movedim(axis, -1)
    N = x.shape[-1]
    M2 = 2 * _next_fast_len(N)
    centered = x - x.mean(dim=-1, keepdim=True)
    f = torch.fft.rfft(centered, n=M2, dim=-1)
    gram = f.real**2 + f.imag**2
    ac = torch.fft.irfft(gram, n=M2, dim=-1)[..., :N]
    ac = ac / torch.arange(N, 0, -1, dtype=ac.dtype, device=ac.device)
    ac = ac / ac[..., :1]
    return ac.movedim(-1, axis)


def autocovariance(x, axis=0):
    x = torch.as_tensor(x).double()
    return autocorrelation(x, axis) * x.var(dim=axis, unbiased=False, keepdim=True)


def _chain_variance_stable(x):
    chain_var = x.var(dim=1, unbiased=True)
    var_within = chain_var.mean(dim=0)
    var_estimator = var_within * (x.shape[1] - 1) / x.shape[1]
    if x.shape[0] > 1:
        var_between = x.mean(dim=1).var(dim=0, unbiased=True)
        var_estimator = var_estimator + var_between
    else:
        var_within = var_estimator
    return var_within, var_estimator


def gelman_rubin(x):
    x = torch.as_tensor(x).double()
    vw, ve = _chain_variance_stable(x)
    return torch.sqrt(ve / vw)


def split_gelman_rubin(x):
    x = torch.as_tensor(x).double()
    h = x.shape[1] // 2
    return gelman_rubin(torch.cat([x[:, :h], x[:, -h:]], dim=0))


def effective_sample_size(x):
    """x: [chains, draws, ...] -> n_eff[...]"""
    x = torch.as_tensor(x).double()
    assert x.dim() >= 2 and x.shape[1] >= 2
    gamma_k_c = autocovariance(x, axis=1)
    vw, ve = _chain_variance_stable(x)
    rho_k = 1.0 - (vw - gamma_k_c.mean(dim=0)) / ve
    rho_k[0] = 1.0
    Rho_k = rho_k[:-1:2] + rho_k[1::2]
    rest = torch.cummin(Rho_k[1:].clamp(min=0), dim=0).values
    Rho_k = torch.cat([Rho_k[:1], rest], dim=0)
    tau = -1.0 + 2.0 * Rho_k.sum(dim=0)
    return (x.shape[0] * x.shape[1]) / tau


def summary(samples, prob=0.9, group_by_chain=True):
    """samples: dict site -> [chains, draws, ...].  Returns dict site -> dict of statistic tensors."""
    out = OrderedDict()
    for name, v in samples.items():
        v = torch.as_tensor(v)
        if not group_by_chain:
            v = v.unsqueeze(0)
        vd = v.double()
        flat = vd.reshape(-1, *vd.shape[2:])
        lo, hi = (1 - prob) / 2, 1 - (1 - prob) / 2
        q = torch.quantile(flat, torch.tensor([lo, 0.5, hi], dtype=flat.dtype, device=flat.device), dim=0) \
            if flat.shape[0] <= 2**24 else None
        out[name] = OrderedDict(
            mean=flat.mean(dim=0),
            std=flat.std(dim=0, unbiased=True),
            median=q[1] if q is not None else flat.median(dim=0).values,
            **({f"{100 * lo:.1f}%": q[0], f"{100 * hi:.1f}%": q[2]} if q is not None else {}),
            n_eff=effective_sample_size(vd),
            r_hat=split_gelman_rubin(vd),
        )
    return out


def print_summary(samples, prob=0.9, group_by_chain=True):
    s = summary(samples, prob, group_by_chain)
    cols = None
    rows = []
    for name, st in s.items():
        cols = list(st.keys())
        shape = st["mean"].shape
        n = int(math.prod(shape)) if len(shape) else 1
        for k in range(n):
            label = name if not len(shape) else f"{name}[{','.join(str(int(i)) for i in torch.unravel_index(torch.tensor(k), shape))}]"
            rows.append((label, [float(st[c].reshape(-1)[k]) for c in cols]))
    print(f"{'':>16}" + "".join(f"{c:>10}" for c in cols))
    for label, vals in rows:
        print(f"{label:>16}" + "".join(f"{v:>10.2f}" for v in vals))
    return s
