"""Consumers of `scripts/jax_bridge.py export` files (trajectories of the UNMODIFIED reference under `PRNGKey(seed)`).

JAX / NumPyro are not installable in the build image, so no real export exists yet: the tests below run on a file of the
same format produced by the ORACLE driven by the restated threefry stream (`oracle/jax_random.py`) -- clearly labelled
`versions = "oracle stand-in"` -- which exercises the whole chain `rng_key -> split/normal/uniform -> trajectory` and the
consumer code; any real export dropped at tests/golden/jax_*.npz is picked up by the same tests and then compares the
CUDA kernels and the oracle with the true reference (fp32, identical decisions, 1e-3).
"""
import glob
import os

import numpy as np
import pytest

from oracle import arwmh_numpy as o
from oracle import jax_random as jr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _standin(path, seeds=(0, 1, 2, 3, 4, 5, 6, 7), steps=120, num_warmup=40, d=10):
    """A bridge-format file from the oracle: chain s uses PRNGKey(seed_s) exactly as ARWMH.sample does (arwmh.py:162)."""
    S = len(seeds)
    nrm = np.empty((steps, S, d), np.float32)
    uni = np.empty((steps, S), np.float32)
    for j, sd in enumerate(seeds):
        nrm[:, j], uni[:, j], _ = jr.arwmh_draws(jr.prng_key(sd), d, steps)
    q0 = np.stack([jr.uniform(jr.split(jr.prng_key(sd), 2)[1], (d,), -2.0, 2.0) for sd in seeds]).astype(np.float32)
    pot = o.make_potential("eight_schools")
    st = o.arwmh_init(pot, q0)
    last, coll = o.arwmh_run(st, pot, steps, draws=(nrm, uni), record_accept=True, num_warmup=num_warmup)
    np.savez_compressed(path, q0=q0, normals=nrm, uniforms=uni, z=coll["z"], potential_energy=coll["potential_energy"],
                        accept=coll["accepts"], loc=last.adapt_state.loc, scale=last.adapt_state.scale,
                        log_step_size=last.adapt_state.log_step_size, seeds=np.array(seeds), num_warmup=num_warmup,
                        lr_decay=2 / 3, model="eight_schools", versions="oracle stand-in (NOT the reference)")


def _files(tmp_path):
    real = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "jax_*.npz")))
    if real:
        return real
    p = str(tmp_path / "standin.npz")
    _standin(p)
    return [p]


def test_threefry_restatement_reproduces_the_exported_draws(tmp_path):
    for f in _files(tmp_path):
        d = np.load(f, allow_pickle=False)
        T, S, dim = d["normals"].shape
        for j, seed in enumerate(d["seeds"][:4]):
            nrm, uni, _ = jr.arwmh_draws(jr.prng_key(int(seed)), dim, min(T, 32))
            np.testing.assert_array_equal(uni, d["uniforms"][: len(uni), j])               # pure bit manipulation
            np.testing.assert_allclose(nrm, d["normals"][: len(nrm), j], rtol=0, atol=5e-7)  # erfinv: libm vs XLA log1p


def test_oracle_follows_the_exported_trajectories(tmp_path):
    for f in _files(tmp_path):
        d = np.load(f, allow_pickle=False)
        if str(d["model"]) != "eight_schools":
            continue
        pot = o.make_potential("eight_schools")
        st = o.arwmh_init(pot, d["q0"].astype(np.float32))
        last, coll = o.arwmh_run(st, pot, d["z"].shape[0], draws=(d["normals"], d["uniforms"]), record_accept=True,
                                 num_warmup=int(d["num_warmup"]), lr_decay=float(d["lr_decay"]))
        same = (coll["accepts"] == d["accept"]).all(axis=0)
        assert same.mean() >= 0.9
        err = np.abs(coll["z"] - d["z"]) / (1 + np.abs(d["z"]))
        assert err.max(axis=(0, 2))[same].max() < 1e-3
        np.testing.assert_allclose(last.adapt_state.scale[same], d["scale"][same], rtol=2e-3, atol=2e-3)


@pytest.mark.gpu
def test_cuda_kernels_follow_the_exported_trajectories(tmp_path):
    import torch

    import adaptive_mcmc_b200 as am
    from adaptive_mcmc_b200 import models

    for f in _files(tmp_path):
        d = np.load(f, allow_pickle=False)
        if str(d["model"]) != "eight_schools":
            continue
        T, S, dim = d["normals"].shape
        s = am.ARWMH(models.eight_schools, num_chains=S, init_strategy=am.init_to_value(torch.from_numpy(d["q0"])))
        st = s.init(0, num_warmup=int(d["num_warmup"]), init_params=None)
        coll, last = s.run(st, T, draws=(torch.from_numpy(d["normals"]), torch.from_numpy(d["uniforms"])), record_accept=True)
        acc = coll["accept"].cpu().numpy()
        same = (acc == d["accept"]).all(axis=0)
        assert same.mean() >= 0.85, same.mean()
        zg = np.concatenate([v.cpu().numpy().reshape(T, S, -1) for v in coll["z"].values()], axis=-1)
        err = (np.abs(zg - d["z"]) / (1 + np.abs(d["z"]))).max(axis=(0, 2))
        assert err[same].max() < 1e-3, err[same].max()
        np.testing.assert_allclose(last.adapt_state.scale.cpu().numpy()[same], d["scale"][same], rtol=3e-3, atol=3e-3)
        # ... and from the seeds alone: rng="jax" regenerates the reference's stream on the GPU (every exported chain is a
        # single-chain run under PRNGKey(seed), run_eight_schools_lr_decay.py:44-46), only q0 comes from the file
        keys = np.stack([jr.prng_key(int(sd)) for sd in d["seeds"]])
        s2 = am.ARWMH(models.eight_schools, num_chains=S, rng="jax", init_strategy=am.init_to_value(torch.from_numpy(d["q0"])))
        st2 = s2.init(keys, num_warmup=int(d["num_warmup"]), init_params=None)
        coll2, _ = s2.run(st2, T, record_accept=True)
        same2 = (coll2["accept"].cpu().numpy() == d["accept"]).all(axis=0)
        assert same2.mean() >= 0.85, same2.mean()
        zg2 = np.concatenate([v.cpu().numpy().reshape(T, S, -1) for v in coll2["z"].values()], axis=-1)
        assert (np.abs(zg2 - d["z"]) / (1 + np.abs(d["z"]))).max(axis=(0, 2))[same2].max() < 1e-3
