"""fp32 accept/reject flips against the float64 oracle, explained instead of tolerated.

A Metropolis step accepts iff u < alpha, alpha = min(1, exp(U - U')).  A float32 implementation computes the energies with an
error e_U (relative to |U|: float32 resolves 6e-8 |U|), so its decision differs from the float64 oracle's only when u falls within alpha * e_dU of alpha
(e_dU = error of U - U').  For u ~ U[0,1) that happens with probability 2 * alpha * e_dU per step.  `analyse` measures
e_dU from the run itself (stored energies of the matching prefix of every chain vs the oracle's), then checks

  * every chain's FIRST flip happened at a step whose margin |u - alpha| is inside the band the measured error allows, and
  * the number of flipped chains is within what the sum of 2 * alpha * e_dU over all (step, chain) pairs predicts (reported
    twice: with the mean energy error -- the expectation -- and with its 99.9 % quantile -- the bound that is asserted)

-- so a pass means "the only decisions that differ are the coin-flips that float32 cannot resolve", not "most chains agree".
After its first flip a chain is a different (equally valid) trajectory and is not compared further.
"""
import numpy as np


def oracle_steps(o, ost, pot, nrm, uni, **kw):
    """Run the NumPy float64 oracle step by step; returns dict(alpha, accept, pe, z) with leading axis T."""
    out = dict(alpha=[], accept=[], pe=[], z=[])
    pe0 = np.asarray(ost.potential_energy, np.float64).copy()
    for t in range(nrm.shape[0]):
        ost, alpha, acc = o.arwmh_step(ost, pot, nrm[t].astype(np.float64), uni[t].astype(np.float64), **kw)
        out["alpha"].append(alpha); out["accept"].append(acc); out["pe"].append(ost.potential_energy.copy()); out["z"].append(ost.z.copy())
    res = {k: np.stack(v) for k, v in out.items()}
    res["pe0"] = pe0
    return res, ost


def analyse(acc_gpu, pe_gpu, orc, uni, label=""):
    """acc_gpu [T,C] bool, pe_gpu [T,C] (energy after each step) or None, orc from oracle_steps, uni [T,C].
    Returns a dict with the measured energy error, observed / predicted flips; raises AssertionError if a flip is unexplained."""
    acc_gpu = np.asarray(acc_gpu, bool)
    T, C = acc_gpu.shape
    differ = acc_gpu != orc["accept"]
    flipped = differ.any(axis=0)
    t_first = np.where(flipped, differ.argmax(axis=0), T)  # steps [0, t_first) match
    live = np.arange(T)[:, None] < t_first[None, :]        # (step, chain) pairs still on the oracle's trajectory
    # energy scale of every step: the chain's energy before the step (the proposal's is of the same order when alpha matters)
    u_prev = np.concatenate([orc["pe0"][None], orc["pe"][:-1]], axis=0)
    scale = 1.0 + np.abs(u_prev)
    if pe_gpu is not None:
        rel = (np.abs(np.asarray(pe_gpu, np.float64) - orc["pe"]) / (1.0 + np.abs(orc["pe"])))[live]
        e_rel = float(np.quantile(rel, 0.999)) if rel.size else 0.0
        e_max = float(rel.max()) if rel.size else 0.0
        e_typ = float(rel.mean()) if rel.size else 0.0
    else:
        e_rel = e_max = e_typ = 0.0
    e_rel = max(e_rel, 6e-8)  # never below half a float32 ulp
    e_u, alpha = e_rel, orc["alpha"]
    margin = np.abs(np.asarray(uni, np.float64) - alpha)
    band = np.minimum(alpha, 1.0) * 2.0 * e_rel * scale   # |u - alpha| inside this band <=> float32 cannot resolve the decision
    # predicted number of chains that flip at least once: 1 - prod(1 - 2 band) over the live-or-would-be-live steps
    p_step = np.minimum(2.0 * band, 1.0)
    p_chain = 1.0 - np.prod(1.0 - p_step, axis=0)
    predicted = float(p_chain.sum())                                   # upper estimate: 99.9 % quantile of the energy error
    typical = float((1.0 - np.prod(1.0 - np.minimum(p_step * (e_typ / e_rel), 1.0), axis=0)).sum())  # with the MEAN error
    observed = int(flipped.sum())
    res = dict(label=label, chains=C, steps=T, energy_relerr_q999=e_u, energy_relerr_max=e_max, flips_observed=observed,
               flips_predicted=predicted, flips_expected_typical=typical, flip_rate_per_step=observed / max(1, int(live.sum())))
    for c in np.nonzero(flipped)[0]:
        t = int(t_first[c])
        allowed = 4.0 * max(band[t, c], 2.0 * e_max * scale[t, c] * min(alpha[t, c], 1.0))
        assert margin[t, c] <= allowed, (f"{label}: chain {c} flips at step {t} with |u - alpha| = {margin[t, c]:.3e}, "
                                         f"alpha = {alpha[t, c]:.4f}, but the measured energy error only allows {allowed:.3e}")
    # Poisson slack: observed within 4 sigma (+2) of the prediction, and the prediction itself must be small
    assert observed <= predicted + 4.0 * np.sqrt(predicted) + 2.0, res
    print("flipcheck", res)
    return res, ~flipped
