"""MCMC -- the chain driver the reference scripts use (``numpyro.infer.MCMC`` as called at
python/scripts/run_eight_schools_wasserstein.py:48-52), re-stated for the fused many-chain kernels:

    mcmc = MCMC(sampler, num_warmup=..., num_samples=..., thinning=..., num_chains=...)
    mcmc.run(rng_key, **data, extra_fields=("potential_energy", "adapt_state"))
    mcmc.get_samples(); mcmc.get_extra_fields(); mcmc.print_summary(); mcmc.last_state

Warm-up + sampling is ONE fused launch when only `z` / `potential_energy` are collected; asking
for per-sample `adapt_state` / `mean_accept_prob` / `as_change` / `i` switches to one launch per
kept sample with device-side snapshots (same results, more launches).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import diagnostics
from .kernels.arwmh import ARWMHAdaptState, ChainBatch

_CHEAP = {"potential_energy", "z"}


class MCMC:
    def __init__(self, sampler, *, num_warmup, num_samples, num_chains=1, thinning=1, progress_bar=False,
                 chain_method="vectorized", **unused):
        if int(thinning) < 1 or int(num_warmup) < 0:
            raise ValueError("thinning must be >= 1 and num_warmup >= 0")
        if int(num_samples) < int(thinning):
            raise ValueError("num_samples must be >= thinning (at least one kept sample)")
        self.sampler = sampler
        self.num_warmup = int(num_warmup)
        self.num_samples = int(num_samples)
        self.num_chains = int(num_chains)
        self.thinning = int(thinning)
        self.progress_bar = progress_bar
        self._samples = None
        self._extra = None
        self._last_state = None
        self._args = ()
        self._kwargs = {}

    @property
    def last_state(self):
        return self._last_state

    @property
    def post_warmup_state(self):
        return None

    def run(self, rng_key, *args, extra_fields=(), init_params=None, **kwargs):
        s = self.sampler
        self._args, self._kwargs = args, kwargs
        state = s.init(rng_key, self.num_warmup, init_params, model_args=args, model_kwargs=kwargs,
                       num_chains=self.num_chains)
        pot = s.potential
        batch = s._batch_from_state(state, copy=False)
        extra_fields = tuple(extra_fields)
        # numpyro.util.fori_collect: collection_size = num_samples // thinning, the first
        # (num_samples % thinning) post-warmup steps are skipped
        n_keep = self.num_samples // self.thinning
        skip = self.num_warmup + self.num_samples % self.thinning
        total = self.num_warmup + self.num_samples
        extras = OrderedDict()
        if set(extra_fields) <= _CHEAP:
            raw = s.run_batch(batch, total, thinning=self.thinning, collect_start=skip,
                              collect=("z", "potential_energy"))
            z = raw["z"].permute(2, 0, 1)  # [C, S, d]
            if "potential_energy" in extra_fields:
                extras["potential_energy"] = raw["potential_energy"].t()
        else:
            if skip:
                s.run_batch(batch, skip, collect=())
            zs, pes, locs, scales, lams, maccs, ascs, its = [], [], [], [], [], [], [], []
            for _ in range(n_keep):
                s.run_batch(batch, self.thinning, collect=())
                zs.append(batch.z.t().clone())
                pes.append(batch.pe.clone())
                if "adapt_state" in extra_fields:
                    locs.append(batch.loc.t().clone())
                    scales.append(batch.dense_scale())
                    lams.append(batch.lam.clone())
                maccs.append(batch.macc.clone())
                ascs.append(batch.asc.clone())
                its.append(batch.i)
            z = torch.stack(zs, dim=1)  # [C, S, d]
            for f in extra_fields:
                if f == "potential_energy":
                    extras[f] = torch.stack(pes, dim=1)
                elif f == "adapt_state":
                    if getattr(s, "_adapt_has_step_size", True):
                        extras[f] = ARWMHAdaptState(torch.stack(locs, 1), torch.stack(scales, 1), torch.stack(lams, 1))
                    else:  # ASSS: (loc, scale) only, the sampler's own record type (asss.py:28)
                        extras[f] = s._adapt_record(torch.stack(locs, 1), torch.stack(scales, 1))
                elif f == "mean_accept_prob":
                    extras[f] = torch.stack(maccs, dim=1)
                elif f == "as_change":
                    extras[f] = torch.stack(ascs, dim=1)
                elif f == "i":
                    extras[f] = torch.tensor(its).unsqueeze(0).expand(batch.C, -1)
                elif f == "z":
                    pass
                else:
                    raise ValueError(f"unknown extra field {f!r}")
        self._z_unconstrained = pot.unravel(z)  # site -> [C, S, ...]
        self._samples = s.postprocess_fn(args, kwargs)(self._z_unconstrained)
        if "z" in extra_fields:
            extras["z"] = self._z_unconstrained
        self._extra = extras
        self._last_state = s._state_from_batch(batch)
        return self

    @staticmethod
    def _flatten(tree, group_by_chain):
        if group_by_chain:
            return tree
        if isinstance(tree, torch.Tensor):
            return tree.reshape(-1, *tree.shape[2:])
        if isinstance(tree, dict):
            return type(tree)((k, MCMC._flatten(v, False)) for k, v in tree.items())
        if isinstance(tree, tuple) and hasattr(tree, "_fields"):
            return type(tree)(*[MCMC._flatten(v, False) for v in tree])
        return tree

    def get_samples(self, group_by_chain=False):
        """Constrained samples (incl. deterministic sites), [C*S, ...] chain-major or [C, S, ...]."""
        return self._flatten(self._samples, group_by_chain)

    def get_extra_fields(self, group_by_chain=False):
        return self._flatten(self._extra, group_by_chain)

    def print_summary(self, prob=0.9, exclude_deterministic=True):
        sites = self._samples
        if exclude_deterministic:
            sites = OrderedDict((k, v) for k, v in sites.items() if k in self._z_unconstrained)
        return diagnostics.print_summary(sites, prob=prob, group_by_chain=True)
