"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chain sharding, the additivity of the
pooled-adaptation sufficient statistics under all-reduce, and the end-of-run moment merge.  The CUDA
kernels themselves are exercised by the -m gpu tests; here the per-shard statistics come from the
oracle so that the exchange logic is checked without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adaptive_mcmc_b200.parallel import _all_reduce_sum, gather_chain_moments, shard_chains


def test_shard_chains_partition():
    for total in (1, 7, 65536, 1_000_000, 16384):
        for world in (1, 2, 3, 8):
            parts = [shard_chains(total, r, world) for r in range(world)]
            assert sum(c for c, _ in parts) == total
            off = 0
            for c, o in parts:
                assert o == off
                off += c
            counts = [c for c, _ in parts]
            assert max(counts) - min(counts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)  # same stream on every rank: the GLOBAL population
        C, d = 1001, 5
        z = rng.normal(size=(C, d)) * np.arange(1, d + 1) + 3.0
        macc = rng.random(C)
        loc = rng.normal(size=d)
        cnt, off = shard_chains(C)  # picks rank/world from the process group
        zl, ml = z[off : off + cnt], macc[off : off + cnt]
        # per-shard sufficient statistics in the layout of amcmc_pooled_stats
        delta = zl - loc
        ii, jj = np.tril_indices(d)
        stats = np.concatenate([[cnt], delta.sum(0), (delta[:, ii] * delta[:, jj]).sum(0), [ml.sum()]])
        t = torch.from_numpy(stats.copy())
        _all_reduce_sum(t)
        dg = z - loc
        ref = np.concatenate([[C], dg.sum(0), (dg[:, ii] * dg[:, jj]).sum(0), [macc.sum()]])
        ok_stats = np.allclose(t.numpy(), ref, rtol=1e-12)
        # identical Robbins-Monro update on every rank from the reduced statistics
        g = 0.5
        cov = (1 - g) * np.eye(d)
        S = np.zeros((d, d)); S[ii, jj] = t.numpy()[1 + d : 1 + d + len(ii)] / C; S = S + np.tril(S, -1).T
        cov = cov + g * S
        ok_cov = np.allclose(cov, (1 - g) * np.eye(d) + g * (dg.T @ dg) / C, rtol=1e-12)
        n, mean, m2 = gather_chain_moments(torch.from_numpy(zl))
        ok_mom = (int(n) == C and np.allclose(mean.numpy(), z.mean(0), rtol=1e-12)
                  and np.allclose((m2 / (n - 1)).numpy(), z.var(0, ddof=1), rtol=1e-10))
        q.put((rank, bool(ok_stats), bool(ok_cov), bool(ok_mom), cnt, off))
    finally:
        dist.destroy_process_group()


def test_pooled_statistics_allreduce_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    for r in res:
        assert r[1] and r[2] and r[3], r
    assert res[0][4] + res[1][4] == 1001 and res[1][5] == res[0][4]
