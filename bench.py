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
    p.add_argument("--probe-d2h", action="store_true",
                   help="only measure pinned device->host / host->device copy bandwidth of every rank, alone and concurrently")
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


REF_ITERS_PER_STEP = 200  # reference arm: fused iterations per bench step (bounded sample, same chain count as ours)


def _ess_of(kept, chains_total):
    """NumPyro-definition ESS (oracle restatement) of kept draws [S, C', d], scaled from C' to all chains."""
    import numpy as np
    from oracle import arwmh_numpy as onp

    x = np.ascontiguousarray(np.transpose(kept, (1, 0, 2)), np.float64)  # [C', S, d]
    ess = onp.effective_sample_size(x)
    return float(np.min(ess)) * (chains_total / x.shape[0])


def run_reference(args):
    """CPU arm.  Order of preference: (1) the UNMODIFIED reference (JAX + NumPyro, `numpyro.infer.MCMC(ARWMH(model),
    num_chains=C, chain_method="vectorized")`) when `import jax, numpyro` works here or from baseline/_ref
    (scripts/jax_bridge.py); (2) the C restatement oracle/arwmh_oracle.c on all host cores.  Same workload, chain count,
    thinning and metric as our arm; each bench step is a bounded sample (REF_ITERS_PER_STEP fused iterations)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    chains = args.chains or wl["chains"]
    T = REF_ITERS_PER_STEP
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import jax_bridge

    jax_ok, jax_detail = jax_bridge.probe()
    ref_root = jax_bridge.find_reference()
    kind, note, val, el, min_ess = "port", "", None, None, None
    if jax_ok and ref_root:
        try:  # the real thing: warm-up steps untimed (also compiles), then K timed steps as one MCMC.run
            rate, sec, mcmc = jax_bridge.time_reference(ref_root, chains, T * args.steps, args.thinning,
                                                        T * max(args.warmup, 1))
            val, el, kind = rate, sec, "reference"
            zs = mcmc.get_samples(group_by_chain=True)
            flat = np.concatenate([np.asarray(v).reshape(v.shape[0], v.shape[1], -1) for k, v in sorted(zs.items())
                                   if k != "theta"], axis=-1)  # [C, S, d]; tau is constrained here (monotone map: same ESS)
            min_ess = _ess_of(np.transpose(flat[:4096], (1, 0, 2)), chains)
            note = f"unmodified reference through numpyro.infer.MCMC ({jax_detail}), JIT compile excluded"
        except Exception as e:  # fall back to the port, and say why
            note = f"jax present ({jax_detail}) but the reference run failed: {repr(e)[:200]}; "
    if val is None:
        from oracle import arwmh_numpy as onp, c_oracle

        ndt = np.float32 if args.dtype == "f32" else np.float64
        q0 = c_oracle.init_uniform(0, chains, 10, dt=ndt)
        st = onp.arwmh_init(onp.make_potential("eight_schools"), q0)
        # Same adaptation stage as the GPU arm for the chains the ESS is taken from: the GPU arm times the iterations after a
        # warm-up of max(W, 3) x 10,000 iterations (n restarts at 1 there, arwmh.py:181).  Warming up all 65,536 chains on the host
        # would take minutes, so only the first 4,096 -- the ones whose draws enter min_ess_per_sec -- get the full warm-up
        # (untimed); the others start cold, which changes nothing about the cost of their steps.
        nw = 10_000 * max(args.warmup, 3)
        n_ess = min(chains, 4096)
        sub = onp.ARWMHState(0, st.z[:n_ess], st.potential_energy[:n_ess], st.mean_accept_prob[:n_ess],
                             onp.ARWMHAdaptState(st.adapt_state.loc[:n_ess], st.adapt_state.scale[:n_ess], st.adapt_state.log_step_size[:n_ess]),
                             st.as_change[:n_ess], 0)
        sub, _ = c_oracle.arwmh_run(sub, "eight_schools", nw, seed=0, n_threads=cores, num_warmup=nw, collect=False)
        # ESS per chain-step of THIS implementation, from a stretch long enough for the estimator (400 kept draws per chain;
        # the timed steps below keep only 4 per step): untimed, then applied to the timed chain-steps/s
        ess_T = min(10_000 * max(args.steps, 1), 100_000)  # the GPU arm's timed stretch (10,000 iterations per step), capped
        sub2, coll2 = c_oracle.arwmh_run(sub, "eight_schools", ess_T, seed=0, thinning=args.thinning, n_threads=cores, num_warmup=nw)
        ess_per_step = _ess_of(coll2["z"], n_ess) / (n_ess * ess_T)
        z, pe, ma, asc = st.z.copy(), st.potential_energy.copy(), st.mean_accept_prob.copy(), st.as_change.copy()
        loc, sc, lam = st.adapt_state.loc.copy(), st.adapt_state.scale.copy(), st.adapt_state.log_step_size.copy()
        z[:n_ess], pe[:n_ess], ma[:n_ess], asc[:n_ess] = sub.z, sub.potential_energy, sub.mean_accept_prob, sub.as_change
        loc[:n_ess], sc[:n_ess], lam[:n_ess] = sub.adapt_state.loc, sub.adapt_state.scale, sub.adapt_state.log_step_size
        st = onp.ARWMHState(nw, z, pe, ma, onp.ARWMHAdaptState(loc, sc, lam), asc, 0)
        for _ in range(max(args.warmup, 1)):
            st, _ = c_oracle.arwmh_run(st, "eight_schools", T, seed=0, thinning=args.thinning, n_threads=cores, num_warmup=nw)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st, coll = c_oracle.arwmh_run(st, "eight_schools", T, seed=0, thinning=args.thinning, n_threads=cores, num_warmup=nw)
        el = time.perf_counter() - t0
        val = chains * T * args.steps / el
        min_ess = ess_per_step * chains * T * args.steps
        note += (f"reference = JAX/NumPyro, not importable here ({jax_detail}); timed arm is the C restatement "
                 "oracle/arwmh_oracle.c (OpenMP over chains)")
    sample = (f"{chains} chains x {T} fused iterations per step, thinning {args.thinning} (ours: 10,000 per step); min_ess_per_sec = timed chain-steps/s x "
              "the min-ESS per chain-step of 4,096 chains over the GPU arm's timed stretch (10,000 iterations per step, at most 100,000) after its warm-up, untimed")
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
        "config": {"workload": wl["label"], "chains_per_gpu": chains, "fused_iterations_per_step": T,
                   "thinning": args.thinning, "sample": sample},
        "min_ess_per_sec": min_ess / el,
        "cpu_baseline": {"value": val, "unit": "chain-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "jax_probe": {"available": jax_ok, "detail": jax_detail, "reference_checkout": bool(ref_root)},
        "note": note,
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
    if args.probe_d2h:
        res = probe_copies(dev, world, local, quick=False)
        if rank == 0:
            print(json.dumps({"probe": "pinned host <-> device copies", "n_gpus": world, **res}))
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload in ("diamonds", "gaussian_ram"):  # secondary workload on its own (profiling, scaling runs)
        K, W = args.steps, max(args.warmup, 3)
        clocks = ClockSampler(local)  # over the whole workload (warm-up included): these functions time themselves
        if rank == 0:
            clocks.start()
        res = (run_diamonds_tc if args.workload == "diamonds" else run_gaussian_ram)(args, world, rank, dev, K, W)
        if rank == 0:
            res["clocks"] = clocks.stop()
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
            "d2h_GBps_per_rank": (state_bytes + S * (d + 1) * Cn * esz) * K / float(te.item()) / 1e9,
            "copy_probe": probe_copies(dev, world, local, quick=True),
        }
        # launches of the host entry per step: its tapering chunk plan (amcmc_host_chunk_samples, csrc/capi.cu)
        left, n_chunks = T // args.thinning, 0
        while left > 0:
            left -= int(_lib.lib().amcmc_host_chunk_samples(left, args.thinning))
            n_chunks += 1
        launches += K * max(1, n_chunks)

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
                         ("sample_Pnx_std_normal", run_sample_pnx), ("configs0_eight_schools_4x10k", run_config0),
                         ("configs1_diamonds_64x50k", run_config1), ("pooled_multi_rank_check", run_pooled_check)):
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
    # The state stays in registers for all fused iterations, so the HBM figure above is the SURVEY 8(d) "state-equivalent"
    # number (it exceeds 1 by design) and NOT a utilisation.  What binds the kernel is instruction issue: thread
    # instructions per chain-step from the committed ncu capture (profiles/instr.json) x chain-steps/s / 32 lanes against
    # 4 schedulers x SMs x the SM clock sampled during the timed region.
    line["roofline"]["binding"] = "issue -- see roofline_issue; frac > 1 here means the state round trip was removed, not 'over peak'"
    line["roofline_issue"] = _issue_roofline(f"{args.workload}_{args.dtype}", value / world, clk, dev)
    if thin1 is not None:
        line["thinning_1"] = thin1
    if extra is not None:
        line["extra_workloads"] = extra
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        cc, ct = Cn, 2000  # same chain count as the GPU arm, a bounded number of iterations (~10-20 s of CPU work)
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
        "roofline_issue": _issue_roofline("gaussian_ram_f32", rate / world, None, dev),
        "roofline": {"bound": "hbm", "unit": "GB/s", "peak": hbm, "achieved": achieved, "frac": achieved / hbm,
                     "binding": "issue / shared-memory port -- see roofline_issue (the factor is SMEM-resident for all fused steps)",
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


# ------------------------------------------------------------------------------------------------
# host <-> device copy ceiling of the box (explains the end-to-end scaling: every rank streams its thinned samples
# to pinned host memory, all ranks into the same host)
# ------------------------------------------------------------------------------------------------
def probe_copies(dev, world, local, quick=True):
    """Pinned D2H / H2D bandwidth of this rank alone (ranks take turns) and of all ranks at once.  GB/s per rank."""
    import torch
    import torch.distributed as dist

    nbytes = (256 if quick else 1024) << 20
    reps = 3 if quick else 8
    hbuf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    d2h = lambda: hbuf.copy_(dbuf, non_blocking=True)
    h2d = lambda: dbuf.copy_(hbuf, non_blocking=True)
    out = {}
    alone = torch.zeros(world, 2, dtype=torch.float64, device=dev)
    rank = dist.get_rank() if world > 1 else 0
    for r in range(world):  # one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            alone[r, 0], alone[r, 1] = timed(d2h), timed(h2d)
    if world > 1:
        dist.all_reduce(alone)
        dist.barrier()
    both = torch.zeros(world, 2, dtype=torch.float64, device=dev)
    both[rank, 0] = timed(d2h)  # all ranks concurrently
    if world > 1:
        dist.barrier()
    both[rank, 1] = timed(h2d)
    if world > 1:
        dist.all_reduce(both)
    out["d2h_GBps_alone"] = [round(float(v), 2) for v in alone[:, 0]]
    out["h2d_GBps_alone"] = [round(float(v), 2) for v in alone[:, 1]]
    out["d2h_GBps_concurrent"] = [round(float(v), 2) for v in both[:, 0]]
    out["h2d_GBps_concurrent"] = [round(float(v), 2) for v in both[:, 1]]
    out["bytes_per_copy"] = nbytes
    try:
        out["numa_nodes"] = len([n for n in os.listdir("/sys/devices/system/node") if n.startswith("node")])
    except OSError:
        out["numa_nodes"] = None
    return out


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[0] and configs[1]: the reference-style runs through the MCMC driver, CPU port beside them
# ------------------------------------------------------------------------------------------------
def _wall(fn, dev):
    import torch

    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize(dev)
    return out, time.perf_counter() - t0


def run_config0(args, world, rank, dev, K, W):
    """configs[0]: eight_schools, 4 chains x 10k steps (+1k warm-up), the reference's own CPU-runnable case, through
    MCMC(ARWMH(model)).run like python/scripts/run_eight_schools_wasserstein.py:48-52.  Latency-bound on a GPU (4 threads)."""
    import numpy as np
    import torch

    import adaptive_mcmc_b200 as am
    from oracle import arwmh_numpy as onp, c_oracle

    def once(seed):
        mcmc = am.MCMC(am.ARWMH(am.models.eight_schools, device=dev, chain_offset=rank * 4), num_warmup=1000, num_samples=10000, num_chains=4)
        mcmc.run(seed, extra_fields=("potential_energy",))
        return mcmc

    once(0)
    times = []
    for k in range(max(K, 3)):
        mcmc, el = _wall(lambda: once(k + 1), dev)
        times.append(el)
    el = float(np.median(times))
    smp = mcmc.get_samples(group_by_chain=True)
    res = {"workload": "eight_schools-eight_schools_centered (d=10), adaptive Metropolis, 4 chains x (1k warm-up + 10k) steps per GPU "
                       "(BASELINE.json configs[0]) through MCMC(ARWMH(model)).run", "metric": "chain-steps/sec",
           "value": world * 4 * 11000 / el, "unit": "chain-steps/s", "wall_s": el, "timing": "host wall clock around mcmc.run (init + one fused launch + postprocess)",
           "mu_mean": float(smp["mu"].mean()), "tau_mean": float(smp["tau"].mean()), "gpu_launches": 2 * max(K, 3),
           "reference_recorded": {"chain_steps_per_s": 5.6e4, "hardware": "laptop CPU, 1 chain (SURVEY section 6)"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        q0 = c_oracle.init_uniform(0, 4, 10, dt=np.float32)
        st = onp.arwmh_init(onp.make_potential("eight_schools"), q0)
        t0 = time.perf_counter()
        c_oracle.arwmh_run(st, "eight_schools", 11000, seed=0, n_threads=4, num_warmup=1000)
        ct = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 4 * 11000 / ct, "unit": "chain-steps/s", "cores": 4, "kind": "port",
                               "sample": f"the same 4 chains x 11,000 steps, C oracle, {ct * 1e3:.1f} ms"}
    return res


def run_config1(args, world, rank, dev, K, W):
    """configs[1]: diamonds (d=26, N=5000), 64 chains x 50k steps from the reference's own start (U(-2,2), identity
    factor) through the MCMC driver; 64 chains = one CTA per chain on the block kernel.  From that start the chains have
    NOT converged after 50k steps (the reference's diamonds runs use a 10^6-step warm-up, run_diamonds_wasserstein.py:67):
    `last_U` is reported next to the time so that nobody reads the run as a converged posterior sample."""
    import numpy as np
    import torch

    import adaptive_mcmc_b200 as am
    from oracle import arwmh_numpy as onp, c_oracle

    data = am.models.synthetic_diamonds(n=5000, k=25, seed=0)

    def once(seed):
        mcmc = am.MCMC(am.ARWMH(am.models.diamonds, device=dev, chain_offset=rank * 64), num_warmup=0, num_samples=50000, thinning=50, num_chains=64)
        mcmc.run(seed, **data, extra_fields=("potential_energy",))
        return mcmc

    once(0)
    mcmc, el = _wall(lambda: once(1), dev)
    pe = mcmc.get_extra_fields(group_by_chain=True)["potential_energy"]
    res = {"workload": "diamonds-diamonds synthetic (d=26, N=5000), adaptive Metropolis, 64 chains x 50k steps per GPU (BASELINE.json "
                       "configs[1]) through MCMC(ARWMH(model)).run, CTA-per-chain block kernel", "metric": "chain-steps/sec",
           "value": world * 64 * 50000 / el, "unit": "chain-steps/s", "wall_s": el, "timing": "host wall clock around mcmc.run",
           "first_U": float(pe[:, 0].mean()), "last_U": float(pe[:, -1].mean()), "U_at_mode": -3283.0,
           "converged": False, "gpu_launches": 4,
           "note": "unconverged by construction: U(-2,2) start with identity factor; the reference's own runs use 10^6 warm-up steps"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        q0 = c_oracle.init_uniform(1, 64, 26, dt=np.float32)
        st = onp.arwmh_init(onp.make_potential("diamonds", **data), q0)
        Tc = 1000
        t0 = time.perf_counter()
        c_oracle.arwmh_run(st, "diamonds", Tc, seed=1, n_threads=cores, collect=False, **data)
        ct = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 64 * Tc / ct, "unit": "chain-steps/s", "cores": cores, "kind": "port",
                               "sample": f"the same 64 chains x {Tc} of the 50,000 steps, C oracle (OpenMP over chains), {ct:.1f} s"}
    return res


def run_pooled_check(args, world, rank, dev, K, W):
    """Correctness of the NCCL pooled-adaptation path on hardware: a FIXED GLOBAL problem (8,192 diamonds chains, Philox
    streams keyed by the global chain id, 6 pooled windows of 50 steps) sharded over the ranks must give the same shared
    adaptation state as the same global chains on one GPU.  With one rank this prints the single-GPU result; the values
    are comparable across the N = 1, 2, 4, 8 lines of a scaling run (sums are float64, so only the summation order differs)."""
    import numpy as np
    import torch

    import adaptive_mcmc_b200 as am
    from adaptive_mcmc_b200.parallel import PooledARWMH, shard_chains

    total = 8192
    data = am.models.synthetic_diamonds(n=5000, k=25, seed=0)
    X, Y = data["X"], data["Y"]
    Xc = np.column_stack([np.ones(len(Y)), X[:, 1:] - X[:, 1:].mean(0)])
    mode = np.concatenate([np.linalg.lstsq(Xc, Y, rcond=None)[0], [np.log(0.123)]])
    q_all = mode[None] + 0.01 * np.random.default_rng(123).normal(size=(total, 26))
    cnt, off = shard_chains(total, rank, world)
    s = PooledARWMH(am.models.diamonds, num_chains=cnt, pool_every=50, device=dev, chain_offset=off,
                    init_strategy=am.init_to_value(torch.from_numpy(q_all[off:off + cnt])))
    s.init(7, model_kwargs=data)
    s.scale.mul_(0.01)
    s.cov.mul_(1e-4)
    for _ in range(6):
        s.run_window(50, collect=())
    zsum = s.batch.z.double().abs().sum().reshape(1)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(zsum)
    return {"workload": "pooled adaptation, 8,192 GLOBAL diamonds chains sharded over the ranks, 6 windows x 50 steps (tcgen05 kernel + "
                        "statistics + all-reduce + update)", "global_chains": total, "log_step_size": round(float(s.log_step_size), 6),
            "loc0": round(float(s.loc[0]), 6), "trace_cov": float(f"{float(torch.trace(s.cov)):.6e}"),
            "sum_abs_z": float(f"{float(zsum):.8e}"), "mean_accept": round(float(s.stats[-1] / s.stats[0]), 5),
            "expect": "identical across N up to the float64 summation order of the pooled statistics"}


def _issue_roofline(key, rate_per_gpu, clk, dev):
    import torch

    ip = os.path.join(ROOT, "profiles", "instr.json")
    ent = json.load(open(ip)).get(key) if os.path.exists(ip) else None
    if not ent:
        return None
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    mhz = (clk or {}).get("sm_mhz") or (clk or {}).get("sm_max_mhz") or 1965.0
    peak = 4 * sms * mhz * 1e6  # warp-instructions per second, one per scheduler per cycle
    achieved = rate_per_gpu * ent["thread_inst_per_chain_step"] / 32.0
    return {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "warp-inst/s", "frac": achieved / peak,
            "thread_inst_per_chain_step": ent["thread_inst_per_chain_step"], "source": ent["source"],
            "note": "peak = 4 schedulers x SMs x sampled SM clock; instruction count from the ncu capture named in `source`"}


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
