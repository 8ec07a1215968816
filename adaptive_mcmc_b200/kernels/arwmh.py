"""ARWMH -- host-side mirror of the reference sampler kernel ``python/kernels/arwmh.py``.

Same class name, constructor kwargs, method names, state record names and field order as the
reference (``ARWMH``, ``ARWMHState``, ``ARWMHAdaptState``; arwmh.py:15-28, :31-276); the
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of ``include/amcmc.h``.
Differences that come with running many chains on a GPU:

  * every state leaf carries a leading chain axis ``C`` and is a CUDA ``torch.Tensor``;
  * ``model`` is one of ``adaptive_mcmc_b200.models`` (the reference's NumPyro model functions,
    pre-compiled), ``potential_fn`` a bound ``PotentialFn``;
  * ``rng_key`` is an integer seed (or a 2-word key): draws come from Philox4x32-10 keyed by
    (seed, global chain id, iteration), or from caller-supplied arrays (shared-draw parity mode);
  * ``run`` / ``run_batch`` fuse K steps into one launch (the reference reaches the same thing
    through ``lax.fori_loop``); ``sample`` is the K = 1 case.

There is no CPU fallback: without the CUDA library or a CUDA device these calls raise.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict, namedtuple

import numpy as np
import torch

from .. import _lib
from ..models import ModelFamily, PotentialFn

_ARWMHStateBase = namedtuple(
    "ARWMHState",
    [
        "i",  # Iteration
        "z",  # Current Point
        "potential_energy",  # Current potential energy
        "mean_accept_prob",  # Running mean of acceptance probabilities
        "adapt_state",  # Mean & Covariance matrix estimate + log of step size
        "as_change",
        "rng_key",  # Random number generator state
    ],
)


class ARWMHState(_ARWMHStateBase):
    """arwmh.py:15-26.  Instances may carry ``_batch`` (the SoA device buffers they were made from)."""


ARWMHAdaptState = namedtuple("ARWMHAdaptState", ["loc", "scale", "log_step_size"])


# ---- init strategies (numpyro.infer.initialization; arwmh.py:12,44) -------------------------
class init_to_uniform:
    """q0 ~ U(-radius, radius) in unconstrained space (NumPyro default radius 2)."""

    def __init__(self, radius=2.0):
        self.radius = float(radius)


class init_to_value:
    def __init__(self, values):
        self.values = values


def _parse_key(rng_key):
    """int seed, or a 2-word key [k0, k1] (jax.random.PRNGKey(s) == [0, s]) -> 64-bit seed."""
    if isinstance(rng_key, (int, np.integer)):
        return int(rng_key) & 0xFFFFFFFFFFFFFFFF
    a = np.asarray(rng_key.cpu() if isinstance(rng_key, torch.Tensor) else rng_key).astype(np.uint64).ravel()
    if a.size == 1:
        return int(a[0])
    if a.size == 2:
        return (int(a[0]) << 32) | (int(a[1]) & 0xFFFFFFFF)
    raise ValueError("rng_key must be an int seed or a 2-word key")


class ChainBatch:
    """Struct-of-arrays device state for C chains (chain index fastest) -- what the CUDA kernels
    read and write in place.  ``scale`` is the packed lower triangle [d(d+1)/2, C]."""

    def __init__(self, potential: PotentialFn, n_chains: int, seed=0, chain_offset=0, alloc=True):
        self.potential = potential
        self.C = int(n_chains)
        self.d = potential.dim
        self.np_ = self.d * (self.d + 1) // 2
        self.i = 0
        self.seed = int(seed)
        self.chain_offset = int(chain_offset)
        self.jax_keys = None  # int32 [2, C] holding the uint32 words of every chain's JAX key (rng = "jax")
        if alloc:
            kw = dict(dtype=potential.dtype, device=potential.device)
            self.z = torch.empty(self.d, self.C, **kw)
            self.loc = torch.empty(self.d, self.C, **kw)
            self.scale = torch.empty(self.np_, self.C, **kw)
            self.pe = torch.empty(self.C, **kw)
            self.macc = torch.empty(self.C, **kw)
            self.lam = torch.empty(self.C, **kw)
            self.asc = torch.empty(self.C, **kw)

    _FIELDS = ("z", "loc", "scale", "pe", "macc", "lam", "asc")

    def clone(self):
        b = ChainBatch(self.potential, self.C, self.seed, self.chain_offset, alloc=False)
        b.i = self.i
        b.jax_keys = None if self.jax_keys is None else self.jax_keys.clone()
        for f in self._FIELDS:
            setattr(b, f, getattr(self, f).clone())
        return b

    def cstruct(self):
        st = _lib.AmcmcState()
        st.n_chains = self.C
        st.dim = self.d
        st.dtype = _lib.AMCMC_F32 if self.potential.dtype == torch.float32 else _lib.AMCMC_F64
        st.i = self.i
        st.z = self.z.data_ptr()
        st.potential_energy = self.pe.data_ptr()
        st.mean_accept_prob = self.macc.data_ptr()
        st.loc = self.loc.data_ptr()
        st.scale = self.scale.data_ptr()
        st.log_step_size = self.lam.data_ptr()
        st.as_change = self.asc.data_ptr()
        return st

    # ---- dense <-> packed ---------------------------------------------------
    def _tril(self):
        return torch.tril_indices(self.d, self.d, device=self.potential.device)

    def dense_scale(self):
        ii, jj = self._tril()
        out = torch.zeros(self.C, self.d, self.d, dtype=self.scale.dtype, device=self.scale.device)
        out[:, ii, jj] = self.scale.t()
        return out

    def set_dense_scale(self, dense):
        ii, jj = self._tril()
        dense = torch.as_tensor(dense, dtype=self.scale.dtype, device=self.scale.device)
        if dense.dim() == 2:
            dense = dense.unsqueeze(0).expand(self.C, self.d, self.d)
        self.scale.copy_(dense[:, ii, jj].t())

    def to_state(self):
        """-> ARWMHState with chain-leading tensors (z as the site dict, scale dense [C,d,d])."""
        pot = self.potential
        z = pot.unravel(self.z.t())
        adapt = ARWMHAdaptState(self.loc.t(), self.dense_scale(), self.lam)
        if self.jax_keys is not None:  # rng = "jax": ARWMHState.rng_key is every chain's current JAX key, uint32 words [C, 2]
            key = self.jax_keys.t().to(torch.int64) & 0xFFFFFFFF
        else:
            key = torch.tensor([self.seed, self.chain_offset], dtype=torch.int64)
        st = ARWMHState(self.i, z, self.pe, self.macc, adapt, self.asc, key)
        st._batch = self
        return st

    @staticmethod
    def from_state(potential, state, copy=True):
        b = getattr(state, "_batch", None)
        if b is not None and b.potential is potential and b.i == int(state.i):
            return b.clone() if copy else b
        zf = potential.ravel(state.z)
        C_ = zf.shape[0]
        rk = state.rng_key.cpu().numpy() if isinstance(state.rng_key, torch.Tensor) else (np.asarray(state.rng_key) if state.rng_key is not None else np.array([0, 0]))
        if rk.ndim == 2 and rk.shape == (C_, 2) and C_ != 1:  # per-chain JAX keys (rng = "jax")
            b = ChainBatch(potential, C_, 0, 0)
            b.jax_keys = torch.from_numpy(np.ascontiguousarray(rk.astype(np.uint64).astype(np.uint32).T).view(np.int32)).to(potential.device)
        else:
            key = rk.ravel()
            b = ChainBatch(potential, C_, int(key[0]), int(key[1]) if key.size > 1 else 0)
        b.i = int(state.i)
        kw = dict(dtype=potential.dtype, device=potential.device)

        def vec(v):
            return torch.as_tensor(v, **kw).reshape(-1).expand(C_).contiguous()

        b.z.copy_(zf.t())
        b.pe.copy_(vec(state.potential_energy))
        b.macc.copy_(vec(state.mean_accept_prob))
        b.asc.copy_(vec(state.as_change))
        loc, scale, lss = state.adapt_state
        loc = torch.as_tensor(loc, **kw)
        if loc.dim() == 1:
            loc = loc.unsqueeze(0).expand(C_, -1)
        b.loc.copy_(loc.t())
        b.set_dense_scale(scale)
        b.lam.copy_(vec(lss))
        return b


class ARWMH:
    """
    ARWMH kernel for adaptive random walk-based Markov Chain Monte Carlo (reference:
    python/kernels/arwmh.py:31-276), many chains at once on one B200.

    Attributes
    ----------
    sample_field : str
        The field name in `ARWMHState` that contains the current sample.
    """

    sample_field = "z"

    def __init__(
        self,
        model=None,
        potential_fn=None,
        lr_decay=2 / 3,
        target_accept_prob=0.234,
        eps=1e-6,
        init_strategy=init_to_uniform,
        *,
        num_chains=1,
        dtype=torch.float32,
        device=None,
        chain_offset=0,
        rng="philox",
    ):
        # arwmh.py:69-70
        if not (model is None) ^ (potential_fn is None):
            raise ValueError("Only one of `model` or `potential_fn` must be specified.")
        if model is not None and not isinstance(model, ModelFamily):
            raise TypeError(
                "model must be one of adaptive_mcmc_b200.models (the reference's NumPyro model functions are "
                "pre-compiled into the CUDA kernels); got %r" % (model,)
            )
        if potential_fn is not None and not isinstance(potential_fn, PotentialFn):
            raise TypeError("potential_fn must be a PotentialFn (ModelFamily.bind(...)): arbitrary Python "
                            "callables cannot be inlined into the CUDA kernels")
        self._model = model
        self._potential_fn = potential_fn
        self._lr_decay = float(lr_decay)
        self._target_accept_prob = float(target_accept_prob)
        self._eps = float(eps)
        self._postprocess_fn = None
        self._init_strategy = init_strategy() if isinstance(init_strategy, type) else init_strategy
        self._num_warmup = 0
        self._num_chains = int(num_chains)
        self._dtype = dtype
        self._device = device
        self._chain_offset = int(chain_offset)
        self._bound_key = None
        self.impl = _lib.IMPL_AUTO
        # rng = "philox": in-kernel counter RNG keyed by (seed, global chain id, iteration) -- the fast path.
        # rng = "jax": the reference's own stream (jax.random threefry2x32; arwmh.py:162-165,174): every chain carries a JAX
        # key, its draws are generated on the GPU (amcmc_jax_draws) and consumed through the external-draws mode, so a run
        # started from `PRNGKey(seed)` and the reference's initial position replays the reference's trajectory.
        if rng not in ("philox", "jax"):
            raise ValueError("rng must be 'philox' or 'jax'")
        self._rng = rng

    @property
    def model(self):
        return self._model

    @property
    def potential(self) -> PotentialFn:
        return self._potential_fn

    # ---- init: arwmh.py:84-138 -------------------------------------------------
    def _bind(self, model_args, model_kwargs):
        if self._model is None:
            return
        key = (id(self._model), tuple(id(a) for a in model_args), tuple(sorted((k, id(v)) for k, v in (model_kwargs or {}).items())))
        if self._potential_fn is None or self._bound_key != key:
            self._potential_fn = self._model.bind(*model_args, dtype=self._dtype, device=self._device, **(model_kwargs or {}))
            self._bound_key = key
        pot = self._potential_fn
        self._postprocess_fn = lambda *a, **k: pot.postprocess

    def init(self, rng_key, num_warmup, init_params, model_args=(), model_kwargs=None, num_chains=None):
        """
        Initialize the ARWMH kernel state (arwmh.py:84-138).

        With a `model`: q0 is drawn by `init_strategy` (default U(-2,2) per unconstrained
        coordinate), `init_params` is ignored exactly as in the reference (:115).  With a
        `potential_fn`, `init_params` (site dict or flat [C, d]) is required (:118-119).
        """
        self._num_warmup = int(num_warmup)
        self._bind(tuple(model_args), model_kwargs)
        pot = self._potential_fn
        if self._rng == "jax":  # an int seed, one key, or per-chain keys [C, 2]; the Philox seed (q0 draws only) is the first key
            seed = _parse_key(np.asarray(rng_key.cpu() if isinstance(rng_key, torch.Tensor) else rng_key).reshape(-1)[:2]
                              if not isinstance(rng_key, (int, np.integer)) else rng_key)
        else:
            seed = _parse_key(rng_key)
        use_given = 0
        radius = 2.0
        zf = None
        if self._model is not None:
            strat = self._init_strategy
            if isinstance(strat, init_to_value):
                zf = pot.ravel(strat.values)
            elif isinstance(strat, init_to_uniform):
                radius = strat.radius
            else:
                raise ValueError("init_strategy must be init_to_uniform or init_to_value")
        else:
            if init_params is None or (isinstance(init_params, dict) and not init_params):
                raise ValueError("Valid value of `init_params` must be provided with `potential_fn`.")
            zf = pot.ravel(init_params)
        C_ = int(num_chains or self._num_chains)
        if zf is not None:
            if zf.shape[0] == 1 and C_ > 1:
                zf = zf.expand(C_, -1)
            C_ = zf.shape[0]
            use_given = 1
        batch = ChainBatch(pot, C_, seed, self._chain_offset)
        if zf is not None:
            batch.z.copy_(zf.t())
        st = batch.cstruct()
        with torch.cuda.device(pot.device):
            rc = _lib.lib().amcmc_arwmh_init(
                pot.handle, C.byref(st), seed, self._chain_offset, radius, use_given,
                C.c_void_p(torch.cuda.current_stream().cuda_stream),
            )
        _lib.check(rc, "amcmc_arwmh_init")
        batch.i = 0
        if self._rng == "jax":
            from ..utils import jax_prng

            keys = jax_prng.chain_keys(rng_key.cpu().numpy() if isinstance(rng_key, torch.Tensor) else rng_key, C_)  # [C, 2]
            batch.jax_keys = torch.from_numpy(np.ascontiguousarray(keys.T).view(np.int32)).to(pot.device)
        return self._state_from_batch(batch)

    # ---- the fused run ----------------------------------------------------------
    def run_batch(self, batch: ChainBatch, num_steps, thinning=1, collect_start=0, collect=("z", "potential_energy"),
                  draws=None, record_accept=False, adapt=True, kernel_kind=_lib.KERNEL_ARWMH):
        """Advance `batch` IN PLACE by `num_steps` fused steps (one kernel launch).  Returns the raw
        collection buffers: dict(z=[S,d,C], potential_energy=[S,C], accept=[T,C] uint8)."""
        pot = batch.potential
        T = int(num_steps)
        thinning = int(thinning)
        S = max(0, (T - int(collect_start)) // thinning) if collect else 0
        kw = dict(dtype=pot.dtype, device=pot.device)
        out = {}
        a = _lib.AmcmcRunArgs()
        a.n_steps = T
        a.thinning = thinning
        a.collect_start = int(collect_start)
        a.num_warmup = self._num_warmup
        a.lr_decay = self._lr_decay
        a.target_accept_prob = self._target_accept_prob
        a.eps = self._eps
        a.adapt = 1 if adapt else 0
        a.seed = batch.seed
        a.chain_offset = batch.chain_offset
        a.kernel_kind = kernel_kind
        a.impl = self.impl
        if draws is None and self._rng == "jax" and kernel_kind != _lib.KERNEL_ASSS and T > 0:
            if batch.jax_keys is None:
                raise ValueError("rng='jax' needs a state made by init() of this sampler (it carries the chains' JAX keys)")
            n_bytes = T * (batch.d + 1) * batch.C * (4 if pot.dtype == torch.float32 else 8)
            if n_bytes > (32 << 30):
                raise ValueError(f"rng='jax' stages all draws of a call in HBM ({n_bytes / 2**30:.0f} GiB here): call run in pieces")
            nrm = torch.empty(T, batch.d, batch.C, **kw)
            uni = torch.empty(T, batch.C, **kw)
            with torch.cuda.device(pot.device):
                _lib.check(_lib.lib().amcmc_jax_draws(batch.jax_keys.data_ptr(), batch.C, batch.d, T,
                                                      _lib.AMCMC_F32 if pot.dtype == torch.float32 else _lib.AMCMC_F64,
                                                      nrm.data_ptr(), uni.data_ptr(),
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "amcmc_jax_draws")
            draws = (nrm, uni)
        if draws is not None:
            nrm, uni = draws
            nrm = torch.as_tensor(nrm, **kw).contiguous()
            uni = torch.as_tensor(uni, **kw).contiguous()
            want_n, want_u = self._draw_shapes(T, batch)
            if tuple(nrm.shape) != want_n or tuple(uni.shape) != want_u:
                raise ValueError(f"draws must be normals{want_n} and uniforms{want_u}; got "
                                 f"{tuple(nrm.shape)}, {tuple(uni.shape)}")
            a.rng_mode = _lib.RNG_EXTERNAL
            a.normals = nrm.data_ptr()
            a.uniforms = uni.data_ptr()
        else:
            a.rng_mode = _lib.RNG_PHILOX
        if S and "z" in collect:
            out["z"] = torch.empty(S, batch.d, batch.C, **kw)
            a.out_z = out["z"].data_ptr()
        if S and "potential_energy" in collect:
            out["potential_energy"] = torch.empty(S, batch.C, **kw)
            a.out_potential_energy = out["potential_energy"].data_ptr()
        if record_accept:
            out["accept"] = torch.empty(T, batch.C, dtype=torch.uint8, device=pot.device)
            a.out_accept = out["accept"].data_ptr()
        st = batch.cstruct()
        with torch.cuda.device(pot.device):
            rc = _lib.lib().amcmc_arwmh_run(
                pot.handle, C.byref(st), C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)
            )
        _lib.check(rc, "amcmc_arwmh_run")
        batch.i = int(st.i)
        return out

    def _draw_shapes(self, T, batch):
        """Device layout of the external draws: normals[T, d, C], uniforms[T, C]."""
        return (T, batch.d, batch.C), (T, batch.C)

    def _batch_from_state(self, state, copy=True):
        return ChainBatch.from_state(self._potential_fn, state, copy=copy)

    def _state_from_batch(self, batch):
        return batch.to_state()

    def _draws_to_device_layout(self, draws):
        """chain-major (normals[T,C,d], uniforms[T,C]) -> device layout."""
        pot = self._potential_fn
        nrm, uni = draws
        nrm = torch.as_tensor(nrm, dtype=pot.dtype, device=pot.device).permute(0, 2, 1).contiguous()
        return nrm, uni

    def run(self, state, num_steps, thinning=1, collect_start=0, collect=("z", "potential_energy"), draws=None,
            record_accept=False):
        """K fused ARWMH.sample steps.  Returns (collected, last_state); collected leaves are
        [S, C, ...] like numpyro.util.fori_collect over vectorised chains (draws, if given, are
        (normals[T,C,d], uniforms[T,C]) chain-major like the state)."""
        pot = self._potential_fn
        batch = self._batch_from_state(state, copy=True)
        if draws is not None:
            draws = self._draws_to_device_layout(draws)
        raw = self.run_batch(batch, num_steps, thinning, collect_start, collect, draws, record_accept)
        coll = OrderedDict()
        if "z" in raw:
            coll["z"] = pot.unravel(raw["z"].permute(0, 2, 1))
        if "potential_energy" in raw:
            coll["potential_energy"] = raw["potential_energy"]
        if "accept" in raw:
            coll["accept"] = raw["accept"].bool()
        return coll, self._state_from_batch(batch)

    def sample(self, state, model_args=(), model_kwargs=None):
        """
        Generate the next sample using the adaptive random walk kernel (arwmh.py:140-207):
        one step for every chain, returning a NEW state (the input is left untouched).
        """
        _, new = self.run(state, 1, collect=())
        return new

    def postprocess_fn(self, args, kwargs):
        # arwmh.py:209-212
        if self._postprocess_fn is None:
            return lambda x: x
        return self._postprocess_fn(*args, **kwargs)

    def get_diagnostics_str(self, state):
        """arwmh.py:214-228; with many chains the chain-averages are reported."""
        acc = float(torch.as_tensor(state.mean_accept_prob).float().mean())
        step = float(torch.exp(torch.as_tensor(state.adapt_state.log_step_size).float()).mean())
        return f"Acceptance rate: {acc:.2f}, Step size: {step:.3f}"

    # ---- frozen many-chain kernel: arwmh.py:230-270 ---------------------------------
    def sample_Pnx(self, rng_key, x, adapt_state, n=1, n_samples=1000, jit_inner=True):
        """P^n(x, .): for each of n_points start points run n_samples independent chains for n steps
        with the adaptation state FROZEN; returns the final positions [n_points, n_samples, ...]
        (site dict if x is a dict).  `jit_inner` is accepted for signature parity and ignored."""
        pot = self._potential_fn
        if pot is None:
            raise ValueError("sample_Pnx needs a bound potential: call init() first or construct with potential_fn")
        is_dict = isinstance(x, dict)
        xf = pot.ravel(x)  # [n_points, d]
        P = xf.shape[0]
        Cn = P * int(n_samples)
        batch = ChainBatch(pot, Cn, _parse_key(rng_key), self._chain_offset)
        start = xf.repeat_interleave(int(n_samples), dim=0)  # chain = point*n_samples + sample
        batch.z.copy_(start.t())
        batch.pe.copy_(pot(start))
        loc, scale, lss = adapt_state
        kw = dict(dtype=pot.dtype, device=pot.device)
        loc = torch.as_tensor(loc, **kw).reshape(-1, pot.dim)[-1]
        scale = torch.as_tensor(scale, **kw).reshape(-1, pot.dim, pot.dim)[-1]
        lss = torch.as_tensor(lss, **kw).reshape(-1)[-1]
        batch.loc.copy_(loc.unsqueeze(1).expand(pot.dim, Cn))
        batch.set_dense_scale(scale)
        batch.lam.fill_(float(lss))
        batch.macc.zero_()
        batch.asc.zero_()
        if not self._run_frozen_shared(batch, loc, scale, lss, int(n)):
            self.run_batch(batch, int(n), collect=(), adapt=False)
        out = batch.z.t().reshape(P, int(n_samples), pot.dim)
        return pot.unravel(out) if is_dict else out

    def _run_frozen_shared(self, batch, loc, dense_scale, lss, n):
        """Frozen steps with ONE adaptation state for all chains (what sample_Pnx is): for many fp32 diamonds chains
        this is the tcgen05 shared-state kernel (amcmc_pooled_run -> diamonds_tc_kernel) instead of one CTA per chain.
        Returns False when that path does not apply (the caller then uses the per-chain frozen kernels)."""
        pot = batch.potential
        sms = torch.cuda.get_device_properties(pot.device).multi_processor_count
        if not (pot.family.name == "diamonds" and pot.dtype == torch.float32 and pot.dim == 26
                and batch.C > 2 * sms and self.impl in (_lib.IMPL_AUTO, _lib.IMPL_TENSOR)):
            return False
        d = pot.dim
        ii, jj = torch.tril_indices(d, d, device=pot.device)
        packed = dense_scale[ii, jj].contiguous()
        loc = loc.contiguous()
        lam = lss.reshape(1).contiguous()
        pool = _lib.AmcmcPooled()
        pool.dim, pool.dtype = d, _lib.AMCMC_F32
        pool.loc, pool.scale, pool.log_step_size = loc.data_ptr(), packed.data_ptr(), lam.data_ptr()
        pool.cov, pool.window = 0, 0
        a = _lib.AmcmcRunArgs()
        a.n_steps, a.thinning, a.collect_start, a.num_warmup = n, 1, n, 0
        a.lr_decay, a.target_accept_prob, a.eps = self._lr_decay, self._target_accept_prob, self._eps
        a.adapt, a.seed, a.chain_offset, a.impl = 0, batch.seed, batch.chain_offset, _lib.IMPL_TENSOR
        a.rng_mode = _lib.RNG_PHILOX
        st = batch.cstruct()
        with torch.cuda.device(pot.device):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().amcmc_pooled_run(pot.handle, C.byref(st), C.byref(pool), C.byref(a), stream),
                       "amcmc_pooled_run")
        batch.i = int(st.i)
        return True

    def get_init_adapt_state(self, rng_key, init_params, model_args=(), model_kwargs={}):
        """Return the first adapt state after initialization (arwmh.py:272-276)."""
        num_warmup = 0
        init_state = self.init(rng_key, num_warmup, init_params, model_args, model_kwargs)
        return init_state.adapt_state
