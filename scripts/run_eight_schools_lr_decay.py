"""The reference's learning-rate-decay experiment (python/scripts/run_eight_schools_lr_decay.py:43-76) on the GPU:
kernels {rwm, sss} x lr_decay {1, 2/3, 1/2} x 100 seeds, each 10^6 steps with the ENTIRE sampler state collected on the
log-spaced grid of utils.kernel_utils.collect_states_logscale (460 snapshots).  The 100 seeds of one (kernel, decay) cell
are the 100 chains of one sampler, so a cell is 460 fused launches; the reference runs 600 single-chain jobs of 10^6
jitted steps each.  With --out DIR the states are written per seed in the reference's pickle layout
(DIR/<kernel>/<decay>/run<seed>.pkl, utils.io.save_states)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_mcmc_b200 as am  # noqa: E402
from adaptive_mcmc_b200.utils import io as amio  # noqa: E402

DECAYS = {"1": 1.0, "2_3": 2 / 3, "1_2": 1 / 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--n-pow", type=int, default=6)
    ap.add_argument("--kernels", default="rwm,sss")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    data = dict(y=am.models.eight_schools.Y, sigma=am.models.eight_schools.SIGMA)
    report = {}
    for kernel_str in a.kernels.split(","):
        for decay_str, lr_decay in DECAYS.items():
            cls = am.ARWMH if kernel_str == "rwm" else am.ASSS
            sampler = cls(am.models.eight_schools, lr_decay=lr_decay, num_chains=a.seeds)
            torch.cuda.synchronize()
            t0 = time.time()
            states = am.collect_states_logscale(0, sampler, data, n_pow=a.n_pow)
            torch.cuda.synchronize()
            dt = time.time() - t0
            asc = states.as_change  # [460, seeds]
            report[f"{kernel_str}/{decay_str}"] = {"seconds": round(dt, 3), "snapshots": int(asc.shape[0]),
                                                   "median_as_change_last": float(asc[-1].median()),
                                                   "min_potential_energy": float(states.potential_energy.min())}
            if a.out:
                d = os.path.join(a.out, kernel_str, decay_str)
                os.makedirs(d, exist_ok=True)
                for seed in range(a.seeds):
                    amio.save_states(states, os.path.join(d, f"run{seed}.pkl"), chain=seed)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
