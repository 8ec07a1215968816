"""Generates tests/golden/*.json.  Run in the build container (needs /root/reference for the
diamonds reference draws; everything else is computed from the reference's notebook outputs,
which are transcribed here with their file:line).  The GPU box never runs this."""
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/python"


def load_jax_pickle(path):
    """Unpickle a dict of jax Arrays without JAX: map
    jax._src.array._reconstruct_array(fun, args, arr_state, aval_state) -> fun(*args)."""

    def _reconstruct_array(fun, args, arr_state, aval_state):
        arr = fun(*args)
        try:
            arr.__setstate__(arr_state)
        except Exception:
            pass
        return np.asarray(arr)

    class U(pickle.Unpickler):
        def find_class(self, module, name):
            if module.startswith("jax") and name == "_reconstruct_array":
                return _reconstruct_array
            return super().find_class(module, name)

    with open(path, "rb") as f:
        return U(f).load()


def main():
    out = {}
    prev = os.path.join(HERE, "reference_pins.json")
    if os.path.exists(prev):  # keys transcribed by hand from the notebooks (quality_tables) are kept as they are
        out = json.load(open(prev))
    # (1) eight_schools ARWMH posterior table, python/jupyter/posteriordb_eight-schools.ipynb:L855-865
    out["eight_schools_arwmh_table"] = {
        "source": "python/jupyter/posteriordb_eight-schools.ipynb:L855-865 (50k warmup + 500k samples, thin 50)",
        "sites": ["mu", "tau", "theta_base[0]", "theta_base[1]", "theta_base[2]", "theta_base[3]",
                  "theta_base[4]", "theta_base[5]", "theta_base[6]", "theta_base[7]"],
        "mean": [4.40, 3.63, 0.32, 0.08, -0.09, 0.05, -0.16, -0.09, 0.36, 0.08],
        "std": [3.29, 3.21, 0.99, 0.93, 0.95, 0.94, 0.92, 0.94, 0.94, 0.96],
        "n_eff": [8787.29, 8866.82, 9080.20, 8919.05, 8305.32, 8798.70, 9130.29, 9527.81, 8791.34, 9331.12],
        "min_potential_energy": 40.975197,
    }
    # (2) energy-scale pins
    out["energy_pins"] = {
        "eight_schools_min_U_100x1e6_steps": [40.638832, "python/jupyter/posteriordb_eight-schools.ipynb:L1172"],
        "diamonds_min_U": [-3283.1575, "python/jupyter/posteriordb_diamonds.ipynb:L2084"],
        "kidiq_min_U": [1874.4451, "python/jupyter/posteriordb_kidiq-kidscore.ipynb cell 64"],
    }
    # (3) eight_schools data literal, posteriordb_eight-schools.ipynb:L502-503
    out["eight_schools_data"] = {"y": [28, 8, -3, 7, -1, 1, 18, 12], "sigma": [15, 10, 16, 11, 9, 11, 10, 18]}
    # (4) diamonds posteriordb reference draws -> moments in the flat unconstrained order
    #     [Intercept, b[0..23], log sigma] (python/scripts/eval_diamonds.py:78-87)
    p = os.path.join(REF, "mcmc_runs", "diamonds-example-references.pkl")
    if os.path.exists(p):
        ref = load_jax_pickle(p)
        x = np.concatenate(
            [np.asarray(ref["Intercept"]).reshape(-1, 1), np.asarray(ref["b"]).reshape(-1, 24),
             np.log(np.asarray(ref["sigma"])).reshape(-1, 1)], axis=1).astype(np.float64)
        out["diamonds_reference_draws"] = {
            "source": "python/mcmc_runs/diamonds-example-references.pkl (posteriordb reference draws, 10000 x 26)",
            "n": int(x.shape[0]),
            "mean": x.mean(0).tolist(),
            "std": x.std(0, ddof=1).tolist(),
            "cov_cond": float(np.linalg.cond(np.cov(x.T))),
            # NOT used by the recovery below -- an independent check of the restated model (sigma couples to the weakly
            # identified coefficients only through the N(0,1) prior of b)
            "corr_logsigma_beta": np.corrcoef(np.column_stack([x[:, 25], x[:, :25]]).T)[0, 1:].tolist(),
        }
        # the draws themselves (float32, 10000 x 26 in the eval_diamonds.py:78-87 order): the `y` of the reference's quality
        # metrics (eval_diamonds.py:62-74, 100-104), needed by scripts/eval_diamonds.py / tests/test_gpu_quality.py on the GPU box
        np.savez_compressed(os.path.join(HERE, "diamonds_reference_draws.npz"), y=x.astype(np.float32))
        # (5) a diamonds-equivalent data set: the Gaussian-linear likelihood depends on the data only through
        #     (N, Xc^T Xc, Xc^T Y, sum Y, Y^T Y); recover them from the draws' mean / covariance / E[sigma^2]
        #     (oracle/diamonds_exact.py) -- posteriordb's diamonds.json itself is not in the image.
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        from oracle import diamonds_exact as de

        st = de.recover_stats_from_draws(x, n=5000)
        pm = de.posterior_moments(st)
        out["diamonds_recovered_stats"] = {
            "source": "recovered from diamonds-example-references.pkl by oracle/diamonds_exact.py:recover_stats_from_draws "
                      "(N = 5000 from posteriordb_diamonds.ipynb:L1452); X1 = [1 | Xc], G = X1^T X1, h = X1^T Y",
            "n": 5000, "G": st["G"].tolist(), "h": st["h"].tolist(), "yy": st["yy"],
            "exact_mean": pm["mean"].tolist(), "exact_std": np.sqrt(np.diag(pm["cov"])).tolist(),
        }
    with open(os.path.join(HERE, "reference_pins.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "reference_pins.json"), list(out))


if __name__ == "__main__":
    main()
