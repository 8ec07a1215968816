"""Multi-GPU layer: chain sharding and pooled (cross-chain) adaptation.

Chains are independent units, so the default path shards them across ranks with NO data-path
collective (one process per GPU; RNG streams are keyed by the GLOBAL chain id, so results do not
depend on the number of GPUs).  The optional pooled-adaptation mode (BASELINE.json configs[3], not
in the reference) shares one adaptation state between all chains of all GPUs: every `pool_every`
steps the per-shard sufficient statistics (2 + d + d(d+1)/2 float64) are all-reduced over
NCCL/NVLink and every rank applies the identical Robbins-Monro update + Cholesky on its device.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .kernels.arwmh import ARWMH, ChainBatch, _parse_key, init_to_uniform


def shard_chains(total_chains, rank=None, world_size=None):
    """Contiguous, balanced partition of the global chain ids -> (count, offset) for this rank."""
    if rank is None or world_size is None:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world_size = torch.distributed.get_rank(), torch.distributed.get_world_size()
        else:
            rank, world_size = 0, 1
    per, rem = divmod(int(total_chains), int(world_size))
    count = per + (1 if rank < rem else 0)
    offset = rank * per + min(rank, rem)
    return count, offset


def _all_reduce_sum(t, group=None):
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=group)
    return t


def gather_chain_moments(z_cd, group=None):
    """End-of-run diagnostics exchange: global (count, mean[d], M2[d]) of per-chain positions
    [C_local, d] merged over ranks (sums are additive; O(d) floats)."""
    z = z_cd.double()
    s = torch.cat([torch.tensor([z.shape[0]], dtype=torch.float64, device=z.device), z.sum(0), (z * z).sum(0)])
    _all_reduce_sum(s, group)
    d = z.shape[1]
    n = s[0]
    mean = s[1 : 1 + d] / n
    m2 = s[1 + d :] - n * mean * mean
    return n, mean, m2


class PooledARWMH:
    """Adaptive random-walk Metropolis with ONE adaptation state (loc, scale, log_step_size) shared
    by all chains on all GPUs.  Within a window of `pool_every` steps every chain runs the frozen
    kernel of the reference's sample_Pnx (arwmh.py:230-249); between windows the shared state moves
    by the reference's Robbins-Monro rule (arwmh.py:183-193) applied to the chain-average innovation.
    diamonds in fp32 runs on the tcgen05 tensor-core path."""

    def __init__(self, model=None, potential_fn=None, lr_decay=2 / 3, target_accept_prob=0.234, eps=1e-6,
                 init_strategy=init_to_uniform, *, num_chains, pool_every=100, dtype=torch.float32, device=None,
                 chain_offset=0, process_group=None, impl=_lib.IMPL_AUTO):
        self._inner = ARWMH(model, potential_fn, lr_decay, target_accept_prob, eps, init_strategy,
                            num_chains=num_chains, dtype=dtype, device=device, chain_offset=chain_offset)
        self.pool_every = int(pool_every)
        self.group = process_group
        self.impl = impl
        self.batch = None
        self.window = 0

    @property
    def potential(self):
        return self._inner.potential

    # ---- state ---------------------------------------------------------------------------------
    def init(self, rng_key, model_args=(), model_kwargs=None, init_params=None):
        st = self._inner.init(rng_key, 0, init_params, model_args, model_kwargs)
        pot = self._inner.potential
        self.batch = ChainBatch.from_state(pot, st, copy=False)
        d = pot.dim
        kw = dict(dtype=pot.dtype, device=pot.device)
        # shared state: loc = global mean of the initial positions, scale = I, log_step = 0, cov = I
        s = torch.cat([torch.tensor([self.batch.C], dtype=torch.float64, device=pot.device), self.batch.z.double().sum(1)])
        _all_reduce_sum(s, self.group)
        self.loc = (s[1:] / s[0]).to(pot.dtype).contiguous()
        ii, jj = torch.tril_indices(d, d, device=pot.device)
        self.scale = (ii == jj).to(pot.dtype).contiguous()
        self.log_step_size = torch.zeros(1, **kw)
        self.cov = torch.eye(d, dtype=torch.float64, device=pot.device).contiguous()
        self.stats = torch.zeros(2 + d + d * (d + 1) // 2, dtype=torch.float64, device=pot.device)
        self.window = 0
        return self

    def _cpool(self):
        p = _lib.AmcmcPooled()
        pot = self._inner.potential
        p.dim = pot.dim
        p.dtype = _lib.AMCMC_F32 if pot.dtype == torch.float32 else _lib.AMCMC_F64
        p.loc = self.loc.data_ptr()
        p.scale = self.scale.data_ptr()
        p.log_step_size = self.log_step_size.data_ptr()
        p.cov = self.cov.data_ptr()
        p.window = self.window
        return p

    def dense_scale(self):
        d = self.potential.dim
        ii, jj = torch.tril_indices(d, d, device=self.scale.device)
        out = torch.zeros(d, d, dtype=self.scale.dtype, device=self.scale.device)
        out[ii, jj] = self.scale
        return out

    # ---- one window ------------------------------------------------------------------------------
    def run_window(self, n_steps, thinning=1, collect_start=0, collect=("z", "potential_energy"), draws=None,
                   record_accept=False, adapt=True):
        """`n_steps` frozen steps for every chain (one launch), then (if adapt) statistics -> all-reduce
        -> Robbins-Monro update.  Returns the raw collection buffers like ARWMH.run_batch."""
        b = self.batch
        pot = b.potential
        inner = self._inner
        T = int(n_steps)
        S = max(0, (T - int(collect_start)) // int(thinning)) if collect else 0
        kw = dict(dtype=pot.dtype, device=pot.device)
        a = _lib.AmcmcRunArgs()
        a.n_steps, a.thinning, a.collect_start, a.num_warmup = T, int(thinning), int(collect_start), 0
        a.lr_decay, a.target_accept_prob, a.eps = inner._lr_decay, inner._target_accept_prob, inner._eps
        a.adapt, a.seed, a.chain_offset, a.impl = 0, b.seed, b.chain_offset, self.impl
        out = {}
        if draws is not None:
            nrm = torch.as_tensor(draws[0], **kw).contiguous()
            uni = torch.as_tensor(draws[1], **kw).contiguous()
            assert tuple(nrm.shape) == (T, b.d, b.C) and tuple(uni.shape) == (T, b.C)
            a.rng_mode, a.normals, a.uniforms = _lib.RNG_EXTERNAL, nrm.data_ptr(), uni.data_ptr()
        else:
            a.rng_mode = _lib.RNG_PHILOX
        if S and "z" in collect:
            out["z"] = torch.empty(S, b.d, b.C, **kw)
            a.out_z = out["z"].data_ptr()
        if S and "potential_energy" in collect:
            out["potential_energy"] = torch.empty(S, b.C, **kw)
            a.out_potential_energy = out["potential_energy"].data_ptr()
        if record_accept:
            out["accept"] = torch.empty(T, b.C, dtype=torch.uint8, device=pot.device)
            a.out_accept = out["accept"].data_ptr()
        st = b.cstruct()
        pool = self._cpool()
        L = _lib.lib()
        with torch.cuda.device(pot.device):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.amcmc_pooled_run(pot.handle, C.byref(st), C.byref(pool), C.byref(a), stream), "amcmc_pooled_run")
            b.i = int(st.i)
            if adapt:
                _lib.check(L.amcmc_pooled_stats(C.byref(st), C.byref(pool), self.stats.data_ptr(), stream), "amcmc_pooled_stats")
                _all_reduce_sum(self.stats, self.group)
                _lib.check(L.amcmc_pooled_update(C.byref(pool), self.stats.data_ptr(), inner._lr_decay,
                                                 inner._target_accept_prob, stream), "amcmc_pooled_update")
                self.window = int(pool.window)
        return out

    def run(self, num_steps, thinning=1, collect=("z", "potential_energy")):
        """num_steps steps in windows of pool_every (K); thinning must divide K or be a multiple of K."""
        K = self.pool_every
        if num_steps % K:
            raise ValueError("num_steps must be a multiple of pool_every")
        if not (K % thinning == 0 or thinning % K == 0):
            raise ValueError("thinning must divide pool_every or be a multiple of it")
        zs, pes = [], []
        done = 0
        for _ in range(num_steps // K):
            if K % thinning == 0:
                raw = self.run_window(K, thinning=thinning, collect=collect)
            elif (done + K) % thinning == 0:
                raw = self.run_window(K, thinning=1, collect_start=K - 1, collect=collect)
            else:
                raw = self.run_window(K, collect=())
            done += K
            if "z" in raw:
                zs.append(raw["z"])
            if "potential_energy" in raw:
                pes.append(raw["potential_energy"])
        coll = OrderedDict()
        pot = self.potential
        if zs:
            coll["z"] = pot.unravel(torch.cat(zs, 0).permute(0, 2, 1))
        if pes:
            coll["potential_energy"] = torch.cat(pes, 0)
        return coll
