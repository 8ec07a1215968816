"""torch.ops.amcmc.* -- the C ABI exposed as PyTorch custom operators (`amcmc::arwmh_run`, `amcmc::logdensity`).

The operators wrap `amcmc_arwmh_run` / `amcmc_potential` 1:1 on plain tensors (struct-of-arrays chain state, chain index
fastest, as include/amcmc.h lays it out), so a PyTorch program can drive the fused kernels without the `ARWMH` class:

    handle = potential.handle                                    # amcmc_model* as an int
    U = torch.ops.amcmc.logdensity(handle, q)                    # q [d, n]  ->  U [n]
    out_z, out_pe = torch.ops.amcmc.arwmh_run(handle, z, pe, macc, loc, scale, lam, asc, i, n_steps, thinning, collect_start,
                                              num_warmup, lr_decay, target, eps, adapt, seed, chain_offset, kernel_kind, impl)

`arwmh_run` updates the state tensors IN PLACE (declared through `mutates_args`) and returns the thinned samples
(out_z [S, d, C], out_pe [S, C]).  Registered with `torch.library` (Python side: the library itself stays free of torch
types); importing this module is what registers them.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@torch.library.custom_op("amcmc::logdensity", mutates_args=())
def logdensity(handle: int, q: torch.Tensor) -> torch.Tensor:
    """potential_fn(z) (python/kernels/arwmh.py:121,170) for n points laid out [d, n]."""
    q = q.contiguous()
    out = torch.empty(q.shape[1], dtype=q.dtype, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.lib().amcmc_potential(C.c_void_p(handle), q.shape[1], q.data_ptr(), out.data_ptr(), _stream(q.device)),
                   "amcmc_potential")
    return out


@logdensity.register_fake
def _(handle, q):
    return q.new_empty(q.shape[1])


@torch.library.custom_op("amcmc::arwmh_run", mutates_args=("z", "pe", "macc", "loc", "scale", "lam", "asc"))
def arwmh_run(handle: int, z: torch.Tensor, pe: torch.Tensor, macc: torch.Tensor, loc: torch.Tensor, scale: torch.Tensor,
              lam: torch.Tensor, asc: torch.Tensor, i: int, n_steps: int, thinning: int, collect_start: int, num_warmup: int,
              lr_decay: float, target_accept_prob: float, eps: float, adapt: bool, seed: int, chain_offset: int,
              kernel_kind: int, impl: int) -> tuple[torch.Tensor, torch.Tensor]:
    """ARWMH.sample (arwmh.py:140-207) fused over n_steps steps for all chains; state tensors are updated in place."""
    d, Cn = z.shape
    S = max(0, (n_steps - collect_start) // thinning)
    out_z = torch.empty(S, d, Cn, dtype=z.dtype, device=z.device)
    out_pe = torch.empty(S, Cn, dtype=z.dtype, device=z.device)
    st = _lib.AmcmcState()
    st.n_chains, st.dim, st.i = Cn, d, i
    st.dtype = _lib.AMCMC_F32 if z.dtype == torch.float32 else _lib.AMCMC_F64
    st.z, st.potential_energy, st.mean_accept_prob = z.data_ptr(), pe.data_ptr(), macc.data_ptr()
    st.loc, st.scale, st.log_step_size, st.as_change = loc.data_ptr(), scale.data_ptr(), lam.data_ptr(), asc.data_ptr()
    a = _lib.AmcmcRunArgs()
    a.n_steps, a.thinning, a.collect_start, a.num_warmup = n_steps, thinning, collect_start, num_warmup
    a.lr_decay, a.target_accept_prob, a.eps = lr_decay, target_accept_prob, eps
    a.adapt, a.rng_mode, a.seed, a.chain_offset = int(adapt), _lib.RNG_PHILOX, seed, chain_offset
    a.kernel_kind, a.impl = kernel_kind, impl
    if S:
        a.out_z, a.out_potential_energy = out_z.data_ptr(), out_pe.data_ptr()
    with torch.cuda.device(z.device):
        _lib.check(_lib.lib().amcmc_arwmh_run(C.c_void_p(handle), C.byref(st), C.byref(a), _stream(z.device)), "amcmc_arwmh_run")
    return out_z, out_pe


@arwmh_run.register_fake
def _(handle, z, pe, macc, loc, scale, lam, asc, i, n_steps, thinning, collect_start, num_warmup, lr_decay, target_accept_prob,
      eps, adapt, seed, chain_offset, kernel_kind, impl):
    d, Cn = z.shape
    S = max(0, (n_steps - collect_start) // thinning)
    return z.new_empty(S, d, Cn), z.new_empty(S, Cn)
