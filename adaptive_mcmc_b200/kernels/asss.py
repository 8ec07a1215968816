"""ASSS -- host-side mirror of the reference's adaptive stereographic slice sampler,
``python/kernels/asss.py`` (class ASSS :98-275, state records :17-30).  SURVEY 8f rank 2.

Same constructor kwargs (`model` xor `potential_fn`, `lr_decay`, `eps`, `init_strategy`), method names and
state record names / field order as the reference; the step (stereographic projection, great-circle shrinkage
with at most 50 trials, mean / Cholesky adaptation) runs in the thread-per-chain CUDA kernel
``csrc/asss_small.cuh`` for the small-dimension models (eight_schools, kidiq, N(0, I_d), custom plugins) and
``csrc/asss_block.cuh`` (one CTA per chain) for diamonds and the dense Gaussian family (d <= 32).
Shared-draw parity mode: ``run(state, T, draws=(normals[T,C,d+1], uniforms[T,C,52]))`` with
uniforms[..., 0] = u_t (asss.py:236), [..., 1] = theta_0 / 2pi (:61), [..., 2+k] = k-th shrinkage draw (:84).
"""
from __future__ import annotations

from collections import namedtuple

import torch

from .. import _lib
from .arwmh import ARWMH, ARWMHAdaptState, ARWMHState, ChainBatch, init_to_uniform

_ASSSStateBase = namedtuple("ASSSState", ["i", "z", "potential_energy", "adapt_state", "as_change", "rng_key"])


class ASSSState(_ASSSStateBase):
    """asss.py:17-28.  Instances may carry ``_batch`` (the SoA device buffers they were made from)."""


ASSSAdaptState = namedtuple("ASSSAdaptState", ["loc", "scale"])

N_UNIFORMS = 52


class ASSS(ARWMH):
    """
    Adaptive Stereographic Slice Sampler kernel for MCMC (reference: python/kernels/asss.py:98-275), many chains
    at once on one B200.
    """

    sample_field = "z"
    _adapt_has_step_size = False      # MCMC(extra_fields=("adapt_state",)) returns ASSSAdaptState records
    _adapt_record = ASSSAdaptState

    def __init__(self, model=None, potential_fn=None, lr_decay=2 / 3, eps=1e-6, init_strategy=init_to_uniform, *,
                 num_chains=1, dtype=torch.float32, device=None, chain_offset=0):
        super().__init__(model, potential_fn, lr_decay, 0.234, eps, init_strategy, num_chains=num_chains, dtype=dtype,
                         device=device, chain_offset=chain_offset)

    # ---- state conversion ------------------------------------------------------------------------
    def _state_from_batch(self, batch):
        st = batch.to_state()
        out = ASSSState(st.i, st.z, st.potential_energy, ASSSAdaptState(st.adapt_state.loc, st.adapt_state.scale),
                        st.as_change, st.rng_key)
        out._batch = batch
        return out

    def _batch_from_state(self, state, copy=True):
        b = getattr(state, "_batch", None)
        if b is not None and b.potential is self._potential_fn and b.i == int(state.i):
            return b.clone() if copy else b
        C_ = self._potential_fn.ravel(state.z).shape[0]
        kw = dict(dtype=self._potential_fn.dtype, device=self._potential_fn.device)
        full = ARWMHState(state.i, state.z, state.potential_energy, torch.zeros(C_, **kw),
                          ARWMHAdaptState(state.adapt_state.loc, state.adapt_state.scale, torch.zeros(C_, **kw)),
                          state.as_change, state.rng_key)
        return ChainBatch.from_state(self._potential_fn, full, copy=copy)

    # ---- draws --------------------------------------------------------------------------------------
    def _draw_shapes(self, T, batch):
        return (T, batch.d + 1, batch.C), (T, N_UNIFORMS, batch.C)

    def _draws_to_device_layout(self, draws):
        pot = self._potential_fn
        nrm, uni = draws
        nrm = torch.as_tensor(nrm, dtype=pot.dtype, device=pot.device).permute(0, 2, 1).contiguous()
        uni = torch.as_tensor(uni, dtype=pot.dtype, device=pot.device).permute(0, 2, 1).contiguous()
        return nrm, uni

    def run_batch(self, batch, num_steps, thinning=1, collect_start=0, collect=("z", "potential_energy"), draws=None,
                  record_accept=False, adapt=True, kernel_kind=_lib.KERNEL_ASSS):
        if record_accept:
            raise ValueError("ASSS always moves: there are no accept decisions to record")
        return super().run_batch(batch, num_steps, thinning, collect_start, collect, draws, False, adapt, kernel_kind)

    def mean_shrink_iterations(self, state):
        """Running mean number of shrinkage trials per step (diagnostic; not part of the reference state)."""
        return self._batch_from_state(state, copy=False).macc

    def get_diagnostics_str(self, state):
        """asss.py:276-277 (potential energy averaged over the chains)."""
        return f"Iteration: {int(state.i)}, Potential Energy: {float(torch.as_tensor(state.potential_energy).float().mean()):.2f}"

    def sample_Pnx(self, rng_key, x, adapt_state, n=1, n_samples=1000, jit_inner=True):
        """asss.py:279-315 -- P^n(x, .) of the slice sampler with the adaptation state (loc, scale) FROZEN: for each
        start point `n_samples` independent chains of `n` steps; returns the final positions [n_points, n_samples, ...]."""
        loc, scale = adapt_state[0], adapt_state[1]
        return super().sample_Pnx(rng_key, x, ARWMHAdaptState(loc, scale, torch.tensor(0.0)), n=n, n_samples=n_samples)
