#!/usr/bin/env python
"""jax_bridge.py -- run the UNMODIFIED reference sampler (savelovme/adaptive-mcmc, JAX + NumPyro) and export what the
B200 path needs to be compared with it: per-step draws, trajectories and final states.

The reference is pure Python on JAX/NumPyro, which are not installable in the build image (no wheels, no network), so
this script could not be executed there; it is written to be ONE command the moment `import jax, numpyro` works
(system install, or `baseline/_ref/` on the path) and a checkout of the reference is available:

    python scripts/jax_bridge.py export --reference /path/to/adaptive-mcmc --model eight_schools \\
           --seeds 0 1 2 3 --steps 300 --num-warmup 100 --out tests/golden/jax_eight_schools.npz
    python scripts/jax_bridge.py time   --reference /path/to/adaptive-mcmc --chains 4096 --steps 2000

Nothing from the reference is copied: the sampler class is imported from `<reference>/python/kernels/arwmh.py`, and the
NumPyro model function is lifted at run time from the reference's own script (`python/scripts/run_<model>_lr_decay.py`,
function `model`) with `ast`, because the scripts themselves open a posteriordb checkout at import time.

export: for every seed s, `rng_key = PRNGKey(s)` exactly as `run_kernel` does (run_eight_schools_lr_decay.py:44-46),
`state = ARWMH(model).init(key, num_warmup, None, (), data)`, then `steps` calls of the jitted `ARWMH.sample`
(arwmh.py:140-207).  Before each call the step's draws are reproduced from the state's key with the very same calls the
kernel makes (`split(key, 3)`, `Normal().sample`, `Uniform().sample`, arwmh.py:162-165,174), so the .npz holds
  q0[S,d], normals[T,S,d], uniforms[T,S], z[T,S,d], potential_energy[T,S], accept[T,S], mean_accept_prob[T,S],
  loc[S,d], scale[S,d,d], log_step_size[S], as_change[T,S], versions.
tests/test_reference_bridge.py consumes such a file (when present) by feeding q0 + draws to the CUDA kernels and to the
oracle and comparing decisions/trajectories -- the comparison with the TRUE reference rather than with a restatement.
It also cross-checks `oracle/jax_random.py` (the threefry restatement) against the real `jax.random`.
"""
from __future__ import annotations

import argparse
import ast
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EIGHT_SCHOOLS = dict(y=[28, 8, -3, 7, -1, 1, 18, 12], sigma=[15, 10, 16, 11, 9, 11, 10, 18])  # posteriordb data (nb :L502-503)


def probe():
    """(ok, detail): can the real reference run here?  Adds baseline/_ref (the offline install target) to sys.path."""
    ref_site = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref_site) and ref_site not in sys.path:
        sys.path.insert(0, ref_site)
    try:
        import jax  # noqa: F401
        import numpyro  # noqa: F401
    except Exception as e:  # ModuleNotFoundError in the build image
        return False, f"{type(e).__name__}: {e}"
    import jax
    import numpyro

    return True, f"jax {jax.__version__}, numpyro {numpyro.__version__}"


def find_reference(path=None):
    for cand in (path, os.environ.get("AMCMC_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref", "adaptive-mcmc")):
        if cand and os.path.isfile(os.path.join(cand, "python", "kernels", "arwmh.py")):
            return cand
    return None


def reference_model(ref_root, name):
    """The NumPyro model function of `python/scripts/run_<name>_lr_decay.py`, compiled from the reference's own text."""
    import jax.numpy as jnp
    import numpyro
    import numpyro.distributions as dist
    import numpyro.infer as infer

    script = {"eight_schools": "run_eight_schools_lr_decay.py", "diamonds": "run_diamonds_lr_decay.py",
              "kidiq": "run_kidiq_kidscore_lr_decay.py"}[name]
    path = os.path.join(ref_root, "python", "scripts", script)
    tree = ast.parse(open(path).read(), path)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "model")
    ns = dict(numpyro=numpyro, dist=dist, infer=infer, jnp=jnp)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["model"]


def reference_kernel(ref_root):
    p = os.path.join(ref_root, "python")
    if p not in sys.path:
        sys.path.insert(0, p)
    from kernels.arwmh import ARWMH  # the unmodified reference class

    return ARWMH


def model_data(name, data_json=None):
    import jax.numpy as jnp
    import numpy as np

    if data_json:
        raw = json.load(open(data_json))
        return {k: jnp.array(v) for k, v in raw.items() if isinstance(v, list)}
    if name == "eight_schools":
        return {k: jnp.array(v) for k, v in EIGHT_SCHOOLS.items()}  # int32, as the reference loads them
    if name == "diamonds":  # the synthetic stand-in the CUDA tests use (posteriordb's diamonds.json is not in the image)
        sys.path.insert(0, ROOT)
        from adaptive_mcmc_b200.models import synthetic_diamonds

        d = synthetic_diamonds(n=5000, k=25, seed=0)
        return dict(X=jnp.array(np.asarray(d["X"], np.float32)), Y=jnp.array(np.asarray(d["Y"], np.float32)))
    raise SystemExit(f"--data <posteriordb json> is required for model {name!r}")


def cmd_export(a):
    ok, detail = probe()
    if not ok:
        raise SystemExit(f"jax/numpyro not importable ({detail}); nothing exported")
    ref = find_reference(a.reference)
    if not ref:
        raise SystemExit("reference checkout not found (pass --reference)")
    import jax
    import numpy as np
    import numpyro.distributions as dist
    from jax import random
    from jax.flatten_util import ravel_pytree

    ARWMH = reference_kernel(ref)
    model = reference_model(ref, a.model)
    data = model_data(a.model, a.data)
    out = {}
    per_seed = []
    for seed in a.seeds:
        kernel = ARWMH(model, lr_decay=a.lr_decay)
        state = kernel.init(random.PRNGKey(seed), a.num_warmup, None, (), data)
        step = jax.jit(lambda s: kernel.sample(s, (), data))
        q0 = np.asarray(ravel_pytree(state.z)[0])
        d = q0.size
        rec = dict(q0=q0, normals=[], uniforms=[], z=[], pe=[], acc=[], macc=[], asc=[])
        for _ in range(a.steps):
            _, kp, ka = random.split(state.rng_key, 3)  # arwmh.py:162
            rec["normals"].append(np.asarray(dist.Normal().sample(kp, sample_shape=(d,))))  # :165
            rec["uniforms"].append(float(dist.Uniform().sample(ka)))  # :174
            new = step(state)
            zf = np.asarray(ravel_pytree(new.z)[0])
            rec["acc"].append(bool((zf != np.asarray(ravel_pytree(state.z)[0])).any()))
            rec["z"].append(zf)
            rec["pe"].append(float(new.potential_energy))
            rec["macc"].append(float(new.mean_accept_prob))
            rec["asc"].append(float(new.as_change))
            state = new
        rec["loc"] = np.asarray(state.adapt_state.loc)
        rec["scale"] = np.asarray(state.adapt_state.scale)
        rec["lam"] = float(state.adapt_state.log_step_size)
        per_seed.append(rec)
    st = lambda k: np.stack([np.asarray(r[k]) for r in per_seed], axis=1)  # [T, S, ...]
    out.update(q0=np.stack([r["q0"] for r in per_seed]), normals=st("normals"), uniforms=st("uniforms"), z=st("z"),
               potential_energy=st("pe"), accept=st("acc"), mean_accept_prob=st("macc"), as_change=st("asc"),
               loc=np.stack([r["loc"] for r in per_seed]), scale=np.stack([r["scale"] for r in per_seed]),
               log_step_size=np.array([r["lam"] for r in per_seed]), seeds=np.array(a.seeds),
               num_warmup=a.num_warmup, lr_decay=a.lr_decay, model=a.model, versions=detail,
               threefry_partitionable=bool(jax.config.jax_threefry_partitionable))
    # cross-check the threefry restatement of the oracle against the real library
    sys.path.insert(0, ROOT)
    from oracle import jax_random as jr

    nrm, uni, _ = jr.arwmh_draws(jr.prng_key(a.seeds[0]), out["q0"].shape[1], min(a.steps, 16))
    ok_bits = np.allclose(uni, out["uniforms"][: len(uni), 0], rtol=0, atol=0)
    ok_nrm = np.allclose(nrm, out["normals"][: len(nrm), 0], rtol=0, atol=4e-7)
    out["oracle_threefry_matches"] = bool(ok_bits and ok_nrm)
    np.savez_compressed(a.out, **out)
    print(json.dumps({"wrote": a.out, "seeds": a.seeds, "steps": a.steps, "versions": detail,
                      "oracle_threefry_matches": out["oracle_threefry_matches"]}))


def time_reference(ref_root, chains, steps, thinning, warmup_steps, repeats=1, model_name="eight_schools", data_json=None):
    """chain-steps/s of the unmodified reference through its public API for many chains:
    numpyro.infer.MCMC(ARWMH(model), num_chains=C, chain_method='vectorized', progress_bar=False).run(...), float32,
    XLA-CPU.  The first run compiles; the timed runs reuse the cached executable (JIT excluded)."""
    import jax
    import numpyro.infer as infer
    from jax import random

    ARWMH = reference_kernel(ref_root)
    model = reference_model(ref_root, model_name)
    data = model_data(model_name, data_json)
    mcmc = infer.MCMC(ARWMH(model), num_warmup=warmup_steps, num_samples=steps, thinning=thinning, num_chains=chains,
                      chain_method="vectorized", progress_bar=False)
    mcmc.run(random.PRNGKey(0), **data)  # compile + first run
    jax.block_until_ready(mcmc.get_samples())
    best = None
    for r in range(repeats):
        t0 = time.perf_counter()
        mcmc.run(random.PRNGKey(1 + r), **data)
        jax.block_until_ready(mcmc.get_samples())
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return chains * (steps + warmup_steps) / best, best, mcmc


def cmd_time(a):
    ok, detail = probe()
    if not ok:
        raise SystemExit(f"jax/numpyro not importable ({detail})")
    ref = find_reference(a.reference)
    if not ref:
        raise SystemExit("reference checkout not found (pass --reference)")
    rate, sec, _ = time_reference(ref, a.chains, a.steps, a.thinning, a.num_warmup, model_name=a.model, data_json=a.data)
    print(json.dumps({"impl": "reference", "kind": "reference", "chain_steps_per_s": rate, "seconds": sec,
                      "chains": a.chains, "steps": a.steps, "cores": os.cpu_count(), "versions": detail}))


def main():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = p.add_subparsers(dest="cmd", required=True)
    for name in ("export", "time", "probe"):
        q = sub.add_parser(name)
        q.add_argument("--reference", default=None)
        q.add_argument("--model", default="eight_schools", choices=["eight_schools", "diamonds", "kidiq"])
        q.add_argument("--data", default=None, help="posteriordb data json (diamonds / kidiq)")
        q.add_argument("--steps", type=int, default=300)
        q.add_argument("--num-warmup", type=int, default=100)
        q.add_argument("--lr-decay", type=float, default=2 / 3)
        q.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2, 3])
        q.add_argument("--chains", type=int, default=4096)
        q.add_argument("--thinning", type=int, default=50)
        q.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "jax_bridge.npz"))
    a = p.parse_args()
    if a.cmd == "probe":
        ok, detail = probe()
        print(json.dumps({"jax_available": ok, "detail": detail, "reference": find_reference(a.reference)}))
    elif a.cmd == "export":
        cmd_export(a)
    else:
        cmd_time(a)


if __name__ == "__main__":
    main()
