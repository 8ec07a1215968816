"""Run outputs in the reference's on-disk format.

The reference scripts pickle what the drivers return -- `pickle.dump(states, f)` after
`collect_states_logscale` (python/scripts/run_eight_schools_lr_decay.py:51-66, run_diamonds_lr_decay.py:65-68)
-- and the notebooks read `states.potential_energy`, `states.as_change`, `states.z`
(posteriordb_eight-schools.ipynb cells 38, 44).  `save_states` writes the same thing for ONE chain of a
many-chain GPU run: a namedtuple pickled under the reference's own class path (`kernels.arwmh.ARWMHState`,
`kernels.arwmh.ARWMHAdaptState`, or the asss equivalents) with NumPy leaves whose leading axis is the collected
sample, so `pickle.load` inside the reference environment yields the reference's own record types."""
from __future__ import annotations

import contextlib
import pickle
import sys
import types
from collections import OrderedDict, namedtuple

import numpy as np
import torch

_REF_CLASSES = {
    "ARWMHState": ("kernels.arwmh", ["i", "z", "potential_energy", "mean_accept_prob", "adapt_state", "as_change", "rng_key"]),
    "ARWMHAdaptState": ("kernels.arwmh", ["loc", "scale", "log_step_size"]),
    "ASSSState": ("kernels.asss", ["i", "z", "potential_energy", "adapt_state", "as_change", "rng_key"]),
    "ASSSAdaptState": ("kernels.asss", ["loc", "scale"]),
}


def _leaf(v, chain):
    """tensor/array with axes [sample, chain, ...] -> numpy [sample, ...] for one chain."""
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    v = np.asarray(v)
    if chain is not None and v.ndim >= 2:
        v = v[:, chain]
    return v


def to_reference_tree(tree, chain=0, _classes=None):
    """Convert a state pytree (namedtuples / dicts of tensors) into plain-NumPy records of one chain.  Namedtuple
    types are re-created with the reference's module path so that pickling refers to `kernels.arwmh.*`."""
    if _classes is None:
        _classes = {}
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        name = type(tree).__name__
        if name in _REF_CLASSES:
            if name not in _classes:
                mod, fields = _REF_CLASSES[name]
                cls = namedtuple(name, fields)
                cls.__module__ = mod
                _classes[name] = cls
            cls = _classes[name]
        else:
            cls = type(tree)
        return cls(*[to_reference_tree(v, chain, _classes) for v in tree])
    if isinstance(tree, dict):
        return {k: to_reference_tree(v, chain, _classes) for k, v in tree.items()}  # plain dict like the reference's z
    if isinstance(tree, (torch.Tensor, np.ndarray)):
        return _leaf(tree, chain)
    return tree


@contextlib.contextmanager
def _reference_modules(classes):
    """Make `kernels.arwmh` / `kernels.asss` importable with OUR record classes while pickling (pickle stores classes
    by reference and verifies the reference resolves)."""
    saved = {}
    try:
        for cls in classes.values():
            parts = cls.__module__.split(".")
            for k in range(1, len(parts) + 1):
                name = ".".join(parts[:k])
                if name not in saved:
                    saved[name] = sys.modules.get(name)
                    if not isinstance(sys.modules.get(name), types.ModuleType) or getattr(sys.modules[name], "__amcmc_stub__", False) is False and name not in sys.modules:
                        m = types.ModuleType(name)
                        m.__amcmc_stub__ = True
                        sys.modules[name] = m
            setattr(sys.modules[cls.__module__], cls.__name__, cls)
        yield
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old


def save_states(states, path, chain=0):
    """Pickle one chain of a collected state pytree (output of collect_states_logscale or a snapshot list) in the
    reference's format.  Returns the converted tree."""
    classes = {}
    tree = to_reference_tree(states, chain, classes)
    with _reference_modules(classes):
        with open(path, "wb") as f:
            pickle.dump(tree, f)
    return tree


def save_samples(mcmc, path, chain=None):
    """Pickle `mcmc.get_samples()` / `get_extra_fields()` as plain dicts of NumPy arrays (the reference pickles the
    whole numpyro MCMC object, python/scripts/run_eight_schools_wasserstein.py:53-57; its consumers only call these
    two getters, eval_eight_schools.py:58-62).  chain=None keeps all chains ([C, S, ...])."""
    def conv(t):
        if isinstance(t, tuple) and hasattr(t, "_fields"):
            return OrderedDict((k, conv(v)) for k, v in zip(t._fields, t))
        if isinstance(t, dict):
            return OrderedDict((k, conv(v)) for k, v in t.items())
        a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
        return a if chain is None else a[chain]
    out = dict(samples=conv(mcmc.get_samples(group_by_chain=True)), extra_fields=conv(mcmc.get_extra_fields(group_by_chain=True)),
               num_warmup=mcmc.num_warmup, num_samples=mcmc.num_samples, thinning=mcmc.thinning)
    with open(path, "wb") as f:
        pickle.dump(out, f)
    return out
