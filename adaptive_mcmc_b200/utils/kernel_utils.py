"""Drivers mirrored from the reference's ``python/utils/kernel_utils.py`` (:8-38)."""
from __future__ import annotations

from collections import OrderedDict

import torch

from ..kernels.arwmh import ARWMHAdaptState, ARWMHState, ChainBatch


def ns_logscale(n_pow=6):
    """kernel_utils.py:8-12 -- the log-spaced iteration grid."""
    return torch.cat(
        [
            torch.arange(0 if p < 1 else 10 ** (p - 1), 10**p, 10 ** (max(0, p - 2))) + 10 ** (max(0, p - 2))
            for p in range(n_pow + 1)
        ]
    )


def _leaves_cat(trees):
    first = trees[0]
    if isinstance(first, torch.Tensor):
        return torch.cat(list(trees))
    if isinstance(first, dict):
        return type(first)((k, _leaves_cat([t[k] for t in trees])) for k in first)
    if isinstance(first, tuple) and hasattr(first, "_fields"):
        return type(first)(*[_leaves_cat([getattr(t, f) for t in trees]) for f in first._fields])
    if isinstance(first, (tuple, list)):
        return type(first)(_leaves_cat([t[k] for t in trees]) for k in range(len(first)))
    return torch.cat([torch.as_tensor(t).reshape(1) if torch.as_tensor(t).dim() == 0 else torch.as_tensor(t) for t in trees])


def concat_trees(trees):
    """kernel_utils.py:14-18 -- concatenate a list of state pytrees along the leading axis."""
    return _leaves_cat(list(trees))


def _snapshot(sampler, batch: ChainBatch):
    """Full sampler state (ARWMHState / ASSSState / ...) at the current iteration, every leaf cloned and given a leading
    sample axis of length 1."""
    st = sampler._state_from_batch(batch)

    def lift(v):
        if isinstance(v, torch.Tensor):
            return v.clone().unsqueeze(0) if v.dim() > 0 else v.clone().reshape(1)
        if isinstance(v, dict):
            return type(v)((k, lift(x)) for k, x in v.items())
        if isinstance(v, tuple) and hasattr(v, "_fields"):
            return type(v)(*[lift(x) for x in v])
        return torch.tensor([v])

    out = lift(st)
    if hasattr(out, "rng_key"):  # one key per chain: (seed, global chain id) identifies the chain's Philox stream
        ids = torch.arange(batch.C, dtype=torch.int64) + int(batch.chain_offset)
        out = out._replace(rng_key=torch.stack([torch.full_like(ids, int(batch.seed) & 0x7FFFFFFFFFFFFFFF), ids], dim=1).unsqueeze(0))
    return out


def collect_states_logscale(rng_key, sampler, model_data: dict, n_pow=6):
    """kernel_utils.py:20-38 -- run 10^n_pow steps and collect the ENTIRE sampler state on the
    log-spaced grid ns_logscale(n_pow).  The reference does this with 7 fori_collect segments of
    thinning 10^max(0,p-2); here each collected point is one fused launch of `thinning` steps
    followed by a device-side snapshot.  Returned leaves are [len(grid), C, ...]; `z` stays
    unconstrained as in the reference (:36 is commented out there)."""
    last_state = sampler.init(rng_key, num_warmup=0, init_params={}, model_args=(), model_kwargs=model_data)
    batch = sampler._batch_from_state(last_state, copy=True)
    collections = []
    for p in range(n_pow + 1):
        lower_idx = 0 if p < 1 else 10 ** (p - 1)
        upper_idx = 10**p
        thinning = 10 ** (max(0, p - 2))
        for _ in range((upper_idx - lower_idx) // thinning):
            sampler.run_batch(batch, thinning, collect=())
            collections.append(_snapshot(sampler, batch))
    return concat_trees(collections)
