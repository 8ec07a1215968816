#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native many-chain adaptive Metropolis sampler.

Metric (BASELINE.json): chain-steps/s (+ min-ESS/s) of the fused ARWMH step.  One bench "step"
is ONE pass of the hot path over one batch: a fused launch of `--mcmc-steps` ARWMH iterations
for `--chains` chains per GPU, with thinned sample collection.

  python bench.py [--gpus N --steps K --warmup W]            # our arm
  python bench.py --impl reference [...]                      # CPU arm (oracle port, all host cores)
  torchrun --nproc-per-node N bench.py --gpus N ...           # N > 1 (chains shard, no data-path collective)

Prints ONE JSON line (rank 0).  See the task contract in DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, d, algorithmic HBM bytes per chain-step (2*w*(d(d+1)/2 + 2d + 3)), default chains)
    "eight_schools": dict(d=10, bytes_per_step_f32=624, chains=65536,
                          label="eight_schools_centered d=10, 65,536 independent chains per GPU (BASELINE.json configs[2])"),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="eight_schools", choices=sorted(WORKLOADS) + ["diamonds", "gaussian_ram"])
    p.add_argument("--chains", type=int, default=None, help="chains per GPU")
    p.add_argument("--mcmc-steps", type=int, default=10000, help="fused ARWMH iterations per bench step")
    p.add_argument("--thinning", type=int, default=50, help="reference thins eight_schools by 50")
    p.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip the secondary diamonds tensor-core workload")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference itself needs JAX + NumPyro, not installable here)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(workload, dtype, chains, mcmc_steps, repeats=1, n_threads=0):
    import numpy as np
    from oracle import arwmh_numpy as onp, c_oracle

    ndt = np.float32 if dtype == "f32" else np.float64
    q0 = c_oracle.init_uniform(0, chains, 10, dt=ndt)
    st = onp.arwmh_init(onp.make_potential("eight_schools"), q0)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        st, _ = c_oracle.arwmh_run(st, "eight_schools", mcmc_steps, seed=0, collect=False, n_threads=n_threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return chains * mcmc_steps / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    chains, T = 4096, 2000  # bounded sample: ~2 s of CPU work per bench step on 8 cores
    import numpy as np
    from oracle import arwmh_numpy as onp, c_oracle

    ndt = np.float32 if args.dtype == "f32" else np.float64
    q0 = c_oracle.init_uniform(0, chains, 10, dt=ndt)
    st = onp.arwmh_init(onp.make_potential("eight_schools"), q0)
    for _ in range(max(args.warmup, 1)):
        st, _ = c_oracle.arwmh_run(st, "eight_schools", T, seed=0, thinning=args.thinning, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, _ = c_oracle.arwmh_run(st, "eight_schools", T, seed=0, thinning=args.thinning, n_threads=cores)
    el = time.perf_counter() - t0
    val = chains * T * args.steps / el
    sample = f"{chains} chains x {T} fused iterations per step (C oracle port, OpenMP over chains)"
    line = {
        "impl": "reference",
        "metric": "chain-steps/sec",
        "value": val,
        "unit": "chain-steps/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["label"], "sample": sample, "thinning": args.thinning},
        "cpu_baseline": {"value": val, "unit": "chain-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = JAX/NumPyro (not installable offline); timed arm is the C restatement oracle/arwmh_oracle.c",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def mark(self):
        return len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for ln in self.proc.stdout:
                self.rows.append(ln.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, lo=0, hi=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[max(lo - 1, 0):(hi + 1 if hi is not None else None)] or self.rows[-2:]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import adaptive_mcmc_b200 as am
    from adaptive_mcmc_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload in ("diamonds", "gaussian_ram"):  # secondary workload on its own (profiling, scaling runs)
        K, W = args.steps, max(args.warmup, 3)
        res = (run_diamonds_tc if args.workload == "diamonds" else run_gaussian_ram)(args, world, rank, dev, K, W)
        if rank == 0:
            res.update({"n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True, "scaling": "weak",
                        "vs_baseline": None, "dtype": "bf16x3 split (fp32 accumulate)" if args.workload == "diamonds" else "f32",
                        "data": "synthetic",
                        "config": {"workload": res.pop("workload")}})
            print(json.dumps(res))
        if world > 1:
            dist.destroy_process_group()
        return
    wl = WORKLOADS[args.workload]
    Cn = args.chains or wl["chains"]
    T = args.mcmc_steps
    K, W = args.steps, max(args.warmup, 3)
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    esz = 4 if args.dtype == "f32" else 8
    d = wl["d"]

    # chains shard across ranks: global chain ids [rank*Cn, (rank+1)*Cn) -> no data-path collective
    sampler = am.ARWMH(am.models.eight_schools, num_chains=Cn, dtype=tdt, device=dev, chain_offset=rank * Cn)
    state = sampler.init(0, num_warmup=W * T, init_params=None)
    batch = am.ChainBatch.from_state(sampler.potential, state, copy=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def one_step(collect=True):
        return sampler.run_batch(batch, T, thinning=args.thinning, collect=("z", "potential_energy") if collect else ())

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    raw = None
    for _ in range(W):  # warm-up = the sampler's adaptation warm-up phase (W*T iterations)
        raw = one_step()  # holding the previous result, like the timed loop: both 577 MB sample buffers get allocated here
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    n_ess = min(Cn, 4096)
    # ESS sample of every timed step, allocated up front: no allocator call (cudaMalloc) between the timed steps
    kept_buf = torch.empty(K, T // args.thinning, d, n_ess, dtype=tdt, device=dev)
    kept = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # two more untimed steps after the last allocation: a sample buffer that the caching allocator hands out for the
    # first time costs its kernel ~8 ms of first-touch (31 ms instead of 22.6 ms), and which step gets it is an accident
    for _ in range(2):
        raw = one_step()
    torch.cuda.synchronize()
    mark0 = clocks.mark()
    for k in range(K):
        flush.fill_(k & 0xFF)  # L2 flush between timed iterations (not timed)
        ev[k][0].record()
        raw = one_step()
        ev[k][1].record()
        kept_buf[k].copy_(raw["z"][:, :, :n_ess])
        kept.append(kept_buf[k])
    torch.cuda.synchronize()
    mark1 = clocks.mark()
    if world > 1:
        dist.barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = K
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    clk = clocks.stop(mark0, mark1) if rank == 0 else None

    value = world * Cn * T * K / (ms_max * 1e-3)

    # min-ESS/s over the timed (post-warm-up) region: NumPyro-definition ESS over all chains of this rank
    zs = torch.cat(kept, dim=0).permute(2, 0, 1)  # [C, S, d]
    ess = am.diagnostics.effective_sample_size(zs)
    min_ess = float(ess.min()) * (Cn / zs.shape[0]) * world
    rhat_max = float(am.diagnostics.split_gelman_rubin(zs).max())
    ess_chains, ess_draws = int(zs.shape[0]), int(zs.shape[1])
    min_ess_per_s = min_ess / (ms_max * 1e-3)
    acc = float(batch.macc.mean())

    # ---- end-to-end through the C ABI with HOST buffers (H2D + run + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        host = {f: getattr(batch, f).cpu().pin_memory() for f in batch._FIELDS}
        S = T // args.thinning
        oz = torch.empty(S, d, Cn, dtype=tdt).pin_memory()
        ope = torch.empty(S, Cn, dtype=tdt).pin_memory()
        hs = _lib.AmcmcState()
        hs.n_chains, hs.dim, hs.dtype, hs.i = Cn, d, (0 if args.dtype == "f32" else 1), batch.i
        hs.z, hs.potential_energy, hs.mean_accept_prob = host["z"].data_ptr(), host["pe"].data_ptr(), host["macc"].data_ptr()
        hs.loc, hs.scale, hs.log_step_size = host["loc"].data_ptr(), host["scale"].data_ptr(), host["lam"].data_ptr()
        hs.as_change = host["asc"].data_ptr()
        a = _lib.AmcmcRunArgs()
        a.n_steps, a.thinning, a.collect_start, a.num_warmup = T, args.thinning, 0, W * T
        a.lr_decay, a.target_accept_prob, a.eps = 2 / 3, 0.234, 1e-6
        a.adapt, a.rng_mode, a.seed, a.chain_offset = 1, 0, 0, rank * Cn
        a.out_z, a.out_potential_energy = oz.data_ptr(), ope.data_ptr()
        h = sampler.potential.handle
        L = _lib.lib()
        for _ in range(2):
            _lib.check(L.amcmc_arwmh_run_host(h, C.byref(hs), C.byref(a)), "run_host")
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            _lib.check(L.amcmc_arwmh_run_host(h, C.byref(hs), C.byref(a)), "run_host")
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        te = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        npk = d * (d + 1) // 2
        state_bytes = Cn * esz * (2 * d + npk + 4)
        e2e = {
            "value": world * Cn * T * K / float(te.item()),
            "unit": "chain-steps/s",
            "h2d_bytes_per_step": state_bytes,
            "d2h_bytes_per_step": state_bytes + S * (d + 1) * Cn * esz,
            "api": "amcmc_arwmh_run_host (C ABI, pinned host buffers)",
        }
        S_ = T // args.thinning
        chunk_S = min(S_, (2048 + args.thinning - 1) // args.thinning) or 1
        launches += K * max(1, -(-S_ // chunk_S))

    extra = None
    thin1 = None
    if not args.no_extra:
        # SURVEY 8d config 3 asks for thinning 1 as well: every state is written (44 B per chain-step)
        T1 = 1000
        for _ in range(2):
            sampler.run_batch(batch, T1, thinning=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sampler.run_batch(batch, T1, thinning=1)
        e1.record()
        torch.cuda.synchronize()
        t1 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t1, op=dist.ReduceOp.MAX)
        rate1 = world * Cn * T1 * 3 / (float(t1.item()) * 1e-3)
        thin1 = {"value": rate1, "unit": "chain-steps/s", "fused_iterations_per_launch": T1,
                 "sample_stream_GBps_per_gpu": rate1 / world * (d + 1) * esz / 1e9}
    if not args.no_extra:  # secondary workloads: run on every rank (they all-reduce)
        del kept, zs, flush
        torch.cuda.empty_cache()
        extra = {}
        for name, fn in (("diamonds_tc", run_diamonds_tc), ("diamonds_tc_adaptive", run_diamonds_adaptive),
                         ("gaussian_ram", run_gaussian_ram), ("asss_eight_schools", run_asss),
                         ("sample_Pnx_std_normal", run_sample_pnx)):
            try:
                extra[name] = fn(args, world, rank, dev, max(3, K // 2), 2)
            except Exception as e:  # the headline line must survive a failure of a secondary workload
                extra[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bps = wl["bytes_per_step_f32"] * (esz // 4)
    per_launch_s = (ms_max * 1e-3) / K
    achieved = bps * Cn * T / per_launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{args.workload}_{args.dtype}_T{T}")  # ncu capture of this exact launch shape
    line = {
        "metric": "chain-steps/sec",
        "value": value,
        "unit": "chain-steps/s",
        "n_gpus": world,
        "steps": K,
        "warmup": W,
        "ms_per_step": ms_max / K,
        "ms_each_step_rank0": [round(a.elapsed_time(b), 3) for a, b in ev],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": args.dtype,
        "data": "synthetic",
        "config": {
            "workload": wl["label"],
            "chains_per_gpu": Cn,
            "fused_iterations_per_step": T,
            "thinning": args.thinning,
            "sampler": "ARWMH lr_decay=2/3 target=0.234 eps=1e-6, Philox4x32-10 in-kernel RNG",
            "l2": "flushed between timed iterations (256 MiB fill)",
            "parallelism": f"chains sharded over {world} GPU(s), no data-path collective",
        },
        "min_ess_per_sec": min_ess_per_s,
        "mean_accept_prob": acc,
        "ess_detail": {"chains_used": ess_chains, "draws_per_chain": ess_draws, "split_rhat_max": rhat_max,
                       "definition": "numpyro.diagnostics.effective_sample_size (Geyer initial monotone), unconstrained coords, "
                                     "scaled from chains_used to all chains"},
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": hbm_peak,
            "unit": "GB/s",
            "frac": achieved / hbm_peak,
            "traffic": traffic,
            "note": ("algorithmic bytes = 2*w*(d(d+1)/2+2d+3) per chain-step (state round trip, SURVEY 8d) x chains x fused "
                     "iterations per launch; peak = MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else
                     "peak = fallback 6650 GB/s (of fallback)"),
        },
    }
    if thin1 is not None:
        line["thinning_1"] = thin1
    if extra is not None:
        line["extra_workloads"] = extra
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        cc, ct = 4096, 5000
        rate, el = cpu_oracle_rate(args.workload, args.dtype, cc, ct, repeats=1, n_threads=cores)
        line["cpu_baseline"] = {
            "value": rate, "unit": "chain-steps/s", "cores": cores, "kind": "port",
            "sample": f"{cc} chains x {ct} iterations of the same workload, C oracle (oracle/arwmh_oracle.c), {el:.1f} s",
        }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# diamonds on the tcgen05 tensor cores (BASELINE.json configs[3] per GPU slice: pooled adaptation)
# ------------------------------------------------------------------------------------------------
def run_diamonds_tc(args, world, rank, dev, K, W):
    """One bench step = `windows` pooled windows of 100 MCMC steps each (frozen tcgen05 kernel +
    statistics kernel + all-reduce + on-device Robbins-Monro/Cholesky update)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import adaptive_mcmc_b200 as am
    from adaptive_mcmc_b200.parallel import PooledARWMH

    # 65,536 chains on one GPU is the configuration the tensor-pipe target is quoted on; BASELINE.json configs[3]
    # (2^20 chains over 8 GPUs) is 131,072 per GPU: `--chains 131072`
    Cn, pool_every, windows = (args.chains or 65536), 100, 10
    data = am.models.synthetic_diamonds(n=5000, k=25, seed=0)
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    q0 = mode[None] + 0.01 * np.random.default_rng(rank).normal(size=(Cn, 26))
    s = PooledARWMH(am.models.diamonds, num_chains=Cn, pool_every=pool_every, device=dev, chain_offset=rank * Cn,
                    init_strategy=am.init_to_value(torch.from_numpy(q0)))
    s.init(0, model_kwargs=data)
    s.scale.mul_(0.01)
    s.cov.mul_(1e-4)

    def step(collect=()):
        out = None
        for w in range(windows):
            out = s.run_window(pool_every, thinning=pool_every, collect=collect if w == windows - 1 else ())
        return out

    for _ in range(W):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(K):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        step()
        ev[k][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    T = pool_every * windows
    rate = world * Cn * T * K / (ms * 1e-3)
    # end to end: positions/energies start in pinned host memory each step, the last window's sample
    # (positions + energies) is read back to the host
    hz = s.batch.z.cpu().pin_memory()
    hpe = s.batch.pe.cpu().pin_memory()
    oz = torch.empty(1, 26, Cn).pin_memory()
    ope = torch.empty(1, Cn).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        s.batch.z.copy_(hz, non_blocking=True)
        s.batch.pe.copy_(hpe, non_blocking=True)
        raw = step(collect=("z", "potential_energy"))
        oz.copy_(raw["z"], non_blocking=True)
        ope.copy_(raw["potential_energy"], non_blocking=True)
        hz.copy_(s.batch.z, non_blocking=True)
        hpe.copy_(s.batch.pe, non_blocking=True)
        torch.cuda.synchronize()
    el = time.perf_counter() - t0
    te = torch.tensor([el], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    per_gpu = rate / world
    return {
        "workload": f"diamonds-diamonds synthetic (d=26, N=5000, Kc=24), {Cn:,} chains per GPU, pooled adaptation every 100 steps "
                    "(BASELINE.json configs[3] slice), tcgen05 split-bf16 likelihood",
        "metric": "chain-steps/sec", "value": rate, "unit": "chain-steps/s", "ms_per_step": ms / K,
        "fused_iterations_per_step": T, "windows_per_step": windows,
        "mean_accept_prob": float(s.batch.macc.mean()), "log_step_size": float(s.log_step_size),
        "e2e": {"value": world * Cn * T * K / float(te.item()), "unit": "chain-steps/s",
                "h2d_bytes_per_step": Cn * 27 * 4, "d2h_bytes_per_step": 2 * Cn * 27 * 4, "api": "PooledARWMH.run_window with pinned host staging"},
        "gpu_launches": K * windows * 4,
        "roofline": {
            "bound": "tensor", "unit": "TFLOP/s", "peak": peak,
            "achieved": per_gpu * 240000 * 3 / 1e12, "frac": per_gpu * 240000 * 3 / 1e12 / peak,
            "executed_tflops": per_gpu * 2 * 5120 * 80 / 1e12, "executed_frac": per_gpu * 2 * 5120 * 80 / 1e12 / peak,
            "algorithmic_single_pass_tflops": per_gpu * 240000 / 1e12, "traffic": _traffic("diamonds_tc_window100"),
            "note": "achieved = chain-steps/s x 2*N*Kc (240,000 flop, SURVEY 8d) x 3 split-bf16 passes; executed = incl. K padding 75->80 "
                    "and row padding 5000->5120; peak = MEASURED_PEAKS.json bf16_tflops_sustained (of measured)",
        },
    }


# ------------------------------------------------------------------------------------------------
# diamonds with PER-CHAIN adaptation (the reference's ARWMH.sample for every chain) on the tensor cores
# ------------------------------------------------------------------------------------------------
def run_diamonds_adaptive(args, world, rank, dev, K, W):
    """One bench step = one fused launch of T ARWMH steps of Cn independent, individually adapted chains
    (diamonds_tc_adapt.cu: tcgen05 likelihood + per-chain LDL^T rank-one update fused with the next proposal)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import adaptive_mcmc_b200 as am

    Cn, T, d = (args.chains or 65536), 500, 26
    data = am.models.synthetic_diamonds(n=5000, k=25, seed=0)
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    q0 = mode[None] + 0.004 * np.random.default_rng(rank).normal(size=(Cn, d))
    s = am.ARWMH(am.models.diamonds, num_chains=Cn, init_strategy=am.init_to_value(torch.from_numpy(q0)), device=dev,
                 chain_offset=rank * Cn)
    st = s.init(rank, num_warmup=0, init_params=None, model_kwargs=data)
    b = am.ChainBatch.from_state(s.potential, st, copy=False)
    b.set_dense_scale(torch.eye(d) * 0.002)  # the reference's identity start rejects everything on this posterior
    for _ in range(max(W, 2)):
        s.run_batch(b, T, collect=())
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(K):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        s.run_batch(b, T, collect=())
        ev[k][1].record()
    torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(c) for a, c in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    rate = world * Cn * T * K / (ms * 1e-3)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    hpeak = float(peaks.get("hbm_gbs", 6650.0))
    per_gpu = rate / world
    state_bytes = 2 * 351 * 4 + 2 * 26 * 4 + 3 * 26 * 4  # factor r+w, running mean r+w, position read twice + written once
    return {
        "workload": f"diamonds synthetic (d=26, N=5000, Kc=24), {Cn:,} chains per GPU, every chain adapts its own mean / "
                    "factor / step size (python/kernels/arwmh.py:140-207), tcgen05 split-bf16 likelihood",
        "metric": "chain-steps/sec", "value": rate, "unit": "chain-steps/s", "ms_per_step": ms / K,
        "fused_iterations_per_step": T, "mean_accept_prob": float(b.macc.mean()), "gpu_launches": K * 6,
        "roofline": {
            "bound": "tensor", "unit": "TFLOP/s", "peak": tpeak,
            "achieved": per_gpu * 240000 * 3 / 1e12, "frac": per_gpu * 240000 * 3 / 1e12 / tpeak,
            "state_stream_GBps": per_gpu * state_bytes / 1e9, "state_stream_frac_of_hbm": per_gpu * state_bytes / 1e9 / hpeak,
            "traffic": _traffic("diamonds_tc_adaptive"),
            "note": f"two alternating streams per SM overlap the tensor-core likelihood with the per-chain state pass ({state_bytes} B "
                    "per chain-step, re-streamed from HBM every step at 65,536 chains); bound by SM<->L2 traffic (3.3 MB per step "
                    "and SM), see profiles/r01_diamonds_tc_adaptive.md",
        },
    }


# ------------------------------------------------------------------------------------------------
# RAM on the correlated Gaussian d = 200 (BASELINE.json configs[4])
# ------------------------------------------------------------------------------------------------
def run_gaussian_ram(args, world, rank, dev, K, W):
    import torch
    import torch.distributed as dist

    import adaptive_mcmc_b200 as am

    d, Cn, T = 200, 16384, 400  # 16,384 chains PER GPU (weak scaling); 400 fused steps per launch
    P = am.models.ar1_precision_chol(d, 0.9)
    s = am.RAM(am.models.gaussian, num_chains=Cn, device=dev, chain_offset=rank * Cn,
               init_strategy=am.init_to_value(torch.zeros(Cn, d)))
    st = s.init(0, num_warmup=0, init_params=None, model_kwargs=dict(prec_chol=P))
    b = am.ChainBatch.from_state(s.potential, st, copy=False)
    b.scale.mul_(2.38 / d**0.5 * 0.3)  # start near the working scale so that accepts and rejects both occur
    for _ in range(W):
        s.run_batch(b, T, collect=())
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(K):  # state (1.3 GB) is far larger than L2: no flush needed
        ev[k][0].record()
        s.run_batch(b, T, thinning=T, collect=("z", "potential_energy"))
        ev[k][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(c) for a, c in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    rate = world * Cn * T * K / (ms * 1e-3)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    bps = 2 * 4 * (d * (d + 1) // 2 + 2 * d + 3)  # 164,024 B: K = 1 state round trip
    achieved = rate / world * bps / 1e9
    return {
        "workload": "synthetic correlated Gaussian d=200 (AR(1) rho=0.9), robust adaptive Metropolis, 16,384 chains per GPU "
                    "(BASELINE.json configs[4]); inputs (1.3 GB of state) exceed L2",
        "metric": "chain-steps/sec", "value": rate, "unit": "chain-steps/s", "ms_per_step": ms / K,
        "fused_iterations_per_step": T, "mean_accept_prob": float(b.macc.mean()), "gpu_launches": K,
        "roofline": {"bound": "hbm", "unit": "GB/s", "peak": hbm, "achieved": achieved, "frac": achieved / hbm,
                     "traffic": _traffic(f"gaussian_ram_T{T}"),
                     "note": "algorithmic bytes = 2*4*(d(d+1)/2+2d+3) = 164,024 per chain-step (state round trip, SURVEY 8d); the "
                             "factor stays in shared memory for all fused steps of a launch, so real HBM traffic is ~1/400 of that"},
    }


# ------------------------------------------------------------------------------------------------
# ASSS (python/kernels/asss.py; SURVEY 8f "next" row) on eight_schools
# ------------------------------------------------------------------------------------------------
def run_asss(args, world, rank, dev, K, W):
    import torch
    import torch.distributed as dist

    import adaptive_mcmc_b200 as am

    Cn, T = 65536, 2000
    s = am.ASSS(am.models.eight_schools, num_chains=Cn, device=dev, chain_offset=rank * Cn)
    st = s.init(0, num_warmup=W * T, init_params=None)
    b = s._batch_from_state(st, copy=False)
    for _ in range(W):
        s.run_batch(b, T, thinning=25)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if world > 1:
        dist.barrier()
    for k in range(K):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        s.run_batch(b, T, thinning=25)
        ev[k][1].record()
    torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(c) for a, c in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"workload": "eight_schools d=10, adaptive stereographic slice sampler (python/kernels/asss.py), 65,536 chains per GPU, "
                        "2,000 fused steps per launch, thinning 25 (the reference's ASSS thinning)",
            "metric": "chain-steps/sec", "value": world * Cn * T * K / (ms * 1e-3), "unit": "chain-steps/s", "ms_per_step": ms / K,
            "mean_shrink_iterations": float(b.macc.mean()), "gpu_launches": K}


def run_sample_pnx(args, world, rank, dev, K, W):
    """The reference's own many-chain benchmark (asumptions_check.ipynb cell 17, :L371-386): ARWMH.sample_Pnx on N(0,1),
    d = 1, 100 start points x 10^5 samples x n = 5 steps with a frozen adaptation state -- 2.28 s wall on the author's
    laptop (2.2e7 chain-steps/s).  Timed here as the whole API call (state set-up + one fused launch + result)."""
    import torch
    import torch.distributed as dist

    import adaptive_mcmc_b200 as am

    pot = am.models.std_normal.bind(d=1, device=dev)
    s = am.ARWMH(potential_fn=pot, chain_offset=rank * 10_000_000)
    x = torch.linspace(-3, 3, 100, device=dev)[:, None]
    ast = am.ARWMHAdaptState(torch.zeros(1), torch.eye(1), torch.tensor(0.0))
    for _ in range(max(W, 6)):  # the first calls grow the caching allocator (a dozen 40 MB state arrays)
        out = s.sample_Pnx(0, x, ast, n=5, n_samples=100_000)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(K):
        out = s.sample_Pnx(k, x, ast, n=5, n_samples=100_000)
        m = float(out.mean())  # device -> host read of a result
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    te = torch.tensor([el], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    el = float(te.item())
    return {"workload": "ARWMH.sample_Pnx, N(0,1) d=1, 100 points x 100,000 samples x n=5 frozen steps per call per GPU "
                        "(python/jupyter/asumptions_check.ipynb cell 17)",
            "metric": "chain-steps/sec", "value": world * 100 * 100_000 * 5 * K / el, "unit": "chain-steps/s", "ms_per_step": 1e3 * el / K,
            "timing": "host wall clock around the whole API call", "reference_recorded": {"seconds_per_call": 2.28, "chain_steps_per_s": 2.2e7,
            "hardware": "laptop CPU, JAX vmap (SURVEY section 6)"}, "last_mean": m, "gpu_launches": K * 3}


def _traffic(key):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(tp)).get(key) if os.path.exists(tp) else None


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
