#!/usr/bin/env python
"""Hardware correctness check of the NCCL pooled-adaptation path (missing #4).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        scripts/check_pooled_multi_gpu.py

Every rank runs its shard of a fixed GLOBAL set of diamonds chains under pooled adaptation (all-reduce of the
float64 sufficient statistics over NCCL every window); rank 0 then repeats the SAME global chains alone on its GPU.
Philox streams are keyed by the global chain id, so the two runs must agree: shared (loc, scale, log step size) to
float64-summation-order accuracy, and positions identically for all but a few chains (a 1-ulp difference of the fp32
shared factor can flip an accept decision).  Prints one JSON line; exit code 1 on disagreement."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_mcmc_b200 as am  # noqa: E402
from adaptive_mcmc_b200.parallel import PooledARWMH, shard_chains  # noqa: E402


def run(total, cnt, off, q_all, data, dev, group, windows, K):
    s = PooledARWMH(am.models.diamonds, num_chains=cnt, pool_every=K, device=dev, chain_offset=off, process_group=group,
                    init_strategy=am.init_to_value(torch.from_numpy(q_all[off:off + cnt])))
    s.init(7, model_kwargs=data)
    s.scale.mul_(0.01)
    s.cov.mul_(1e-4)
    for _ in range(windows):
        s.run_window(K, collect=())
    return s


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    total, windows, K = 16384, 8, 50
    data = am.models.synthetic_diamonds(n=5000, k=25, seed=0)
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    q_all = mode[None] + 0.01 * np.random.default_rng(123).normal(size=(total, 26))
    cnt, off = shard_chains(total, rank, world)
    s = run(total, cnt, off, q_all, data, dev, None, windows, K)
    z_parts = [torch.empty(26, shard_chains(total, r, world)[0], device=dev) for r in range(world)]
    if world > 1:
        dist.all_gather(z_parts, s.batch.z.contiguous())
    else:
        z_parts = [s.batch.z]
    ok, out = True, None
    if rank == 0:
        solo_group = dist.new_group([0]) if world > 1 else None
    elif world > 1:
        dist.new_group([0])  # collective call: every rank must take part
    if rank == 0:
        s1 = run(total, total, 0, q_all, data, dev, solo_group, windows, K)
        z_multi = torch.cat(z_parts, dim=1)
        same = (z_multi == s1.batch.z).all(dim=0).float().mean().item()
        d_loc = (s.loc - s1.loc).abs().max().item()
        d_scale = (s.scale - s1.scale).abs().max().item() / s1.scale.abs().max().item()
        d_lam = abs(float(s.log_step_size) - float(s1.log_step_size))
        ok = same > 0.995 and d_loc < 1e-5 and d_scale < 1e-4 and d_lam < 1e-5
        out = {"check": "pooled adaptation N ranks vs the same global chains on 1 GPU", "n_gpus": world, "global_chains": total,
               "windows": windows, "steps_per_window": K, "chains_bitwise_identical": same, "max_abs_dloc": d_loc,
               "max_rel_dscale": d_scale, "abs_dlog_step": d_lam, "log_step_size": float(s1.log_step_size), "ok": ok}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
