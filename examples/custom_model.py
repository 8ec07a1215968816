"""A model that is not one of the built-in families: Bayesian logistic regression, written as CUDA snippets and run
with the same drivers as the reference's models (the reference takes any NumPyro model function, arwmh.py:43-78; here
the potential is compiled into the fused kernels as a plugin, adaptive_mcmc_b200/custom.py).

    python examples/custom_model.py            # needs a B200 and nvcc; ~25 s for the first build of each plugin

Two variants of the same posterior:
  * `potential=`  -> thread-per-chain register kernels (the whole likelihood loop inside one thread; right for a few
                     hundred rows and many chains)
  * `row_term=`   -> CTA-per-chain kernels (rows spread over the 256 threads of the chain's CTA; right for many rows)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_mcmc_b200 as am  # noqa: E402

ROW = '''
    const R eta = q[0] + q[1] * a0[2 * i] + q[2] * a0[2 * i + 1];
    const R sp = eta > (R)0 ? eta + Num<R>::log1p(Num<R>::exp(-eta)) : Num<R>::log1p(Num<R>::exp(eta));
    return sp - a1[i] * eta;                               // -log Bernoulli(y_i | sigmoid(eta))'''
PRIOR = "    return (R)0.02 * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);   // beta ~ N(0, 5^2)"
THREAD = '''
    R u = (R)0.02 * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    for (int64_t i = 0; i < n1; ++i) {
      const R eta = q[0] + q[1] * a0[2 * i] + q[2] * a0[2 * i + 1];
      const R sp = eta > (R)0 ? eta + Num<R>::log1p(Num<R>::exp(-eta)) : Num<R>::log1p(Num<R>::exp(eta));
      u += sp - a1[i] * eta;
    }
    return u;'''


def main():
    rng = np.random.default_rng(0)
    beta = np.array([0.5, 1.2, -0.7])
    for n, kind in ((300, "thread"), (20000, "rows")):
        x = rng.normal(size=(n, 2))
        y = (rng.random(n) < 1 / (1 + np.exp(-(beta[0] + x @ beta[1:])))).astype(np.float64)
        if kind == "thread":
            fam = am.custom_model("example_logistic_thread", [("beta", (3,))], arrays=["x", "y"], potential=THREAD)
            chains = 4096
        else:
            fam = am.custom_model("example_logistic_rows", [("beta", (3,))], arrays=["x", "y"], rows="y", row_term=ROW, prior=PRIOR)
            chains = 256
        mcmc = am.MCMC(am.ARWMH(fam), num_warmup=5000, num_samples=20000, thinning=20, num_chains=chains)
        mcmc.run(0, x=x, y=y)
        b = mcmc.get_samples()["beta"].double()
        print(f"{kind:6s} n = {n:6d}, {chains} chains: posterior mean {b.mean(0).cpu().numpy().round(3)}  sd {b.std(0).cpu().numpy().round(3)}"
              f"  (generating beta {beta})")


if __name__ == "__main__":
    main()
