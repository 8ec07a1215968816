"""Host-side logic that needs no GPU: interface mirror of python/kernels/arwmh.py, drivers, diagnostics."""
import numpy as np
import pytest
import torch

import adaptive_mcmc_b200 as am
from adaptive_mcmc_b200.kernels import arwmh as K
from oracle import arwmh_numpy as o


def test_ctor_contract_matches_reference():
    # arwmh.py:69-70
    with pytest.raises(ValueError):
        am.ARWMH()
    with pytest.raises(ValueError):
        am.ARWMH(model=am.models.eight_schools, potential_fn=object())
    with pytest.raises(TypeError):
        am.ARWMH(model=lambda: None)
    s = am.ARWMH(am.models.eight_schools)
    assert s.sample_field == "z" and s.model is am.models.eight_schools
    assert s._lr_decay == pytest.approx(2 / 3) and s._target_accept_prob == 0.234 and s._eps == 1e-6
    assert K.ARWMHState._fields == ("i", "z", "potential_energy", "mean_accept_prob", "adapt_state", "as_change", "rng_key")
    assert K.ARWMHAdaptState._fields == ("loc", "scale", "log_step_size")
    assert s.postprocess_fn((), {})(5) == 5  # identity before init (arwmh.py:210-211)


def test_parse_key():
    assert K._parse_key(7) == 7
    assert K._parse_key(np.array([0, 7], dtype=np.uint32)) == 7
    assert K._parse_key(torch.tensor([1, 2])) == (1 << 32) | 2
    with pytest.raises(ValueError):
        K._parse_key([1, 2, 3])


def test_ns_logscale_and_concat_trees():
    g = am.ns_logscale(6)
    np.testing.assert_array_equal(g.numpy(), o.ns_logscale(6))
    a = K.ARWMHAdaptState(torch.zeros(2, 3), torch.zeros(2, 3, 3), torch.zeros(2))
    t1 = K.ARWMHState(torch.tensor([1, 2]), {"mu": torch.zeros(2)}, torch.zeros(2), torch.zeros(2), a, torch.zeros(2), torch.zeros(2, 2))
    cat = am.concat_trees([t1, t1, t1])
    assert cat.z["mu"].shape == (6,) and cat.adapt_state.scale.shape == (6, 3, 3) and cat.i.shape == (6,)


def test_diagnostics_match_oracle():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 400, 2)).cumsum(axis=1) * 0.05 + rng.normal(size=(3, 400, 2))
    np.testing.assert_allclose(am.diagnostics.effective_sample_size(torch.from_numpy(x)).numpy(), o.effective_sample_size(x), rtol=1e-9)
    np.testing.assert_allclose(am.diagnostics.split_gelman_rubin(torch.from_numpy(x)).numpy(), o.split_gelman_rubin(x), rtol=1e-12)
    np.testing.assert_allclose(am.diagnostics.gelman_rubin(torch.from_numpy(x)).numpy(), o.gelman_rubin(x), rtol=1e-12)


def test_missing_library_fails_loudly(monkeypatch):
    from adaptive_mcmc_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libamcmc.so")
    with pytest.raises(_lib.AmcmcError):
        _lib.lib()


def test_cpu_device_rejected():
    with pytest.raises((ValueError, RuntimeError, AssertionError)):
        am.models.eight_schools.bind(device="cpu")


def test_save_states_in_reference_pickle_format(tmp_path):
    """a collected state tree pickles under the reference's class paths with NumPy leaves."""
    import pickle, sys, types
    from collections import namedtuple, OrderedDict
    from adaptive_mcmc_b200.utils.io import save_states
    S, C, d = 5, 3, 4
    st = K.ARWMHState(torch.arange(1, S + 1), OrderedDict(mu=torch.randn(S, C), theta=torch.randn(S, C, 2)), torch.randn(S, C),
                      torch.rand(S, C), K.ARWMHAdaptState(torch.randn(S, C, d), torch.randn(S, C, d, d), torch.randn(S, C)),
                      torch.rand(S, C), torch.zeros(S, 2, dtype=torch.int64))
    p = tmp_path / "run0.pkl"
    save_states(st, str(p), chain=1)
    assert "kernels.arwmh" not in sys.modules  # the stub modules are gone after the dump
    # the reference environment: python/kernels/arwmh.py defines the record types
    kern = types.ModuleType("kernels"); arw = types.ModuleType("kernels.arwmh")
    arw.ARWMHState = namedtuple("ARWMHState", K.ARWMHState._fields); arw.ARWMHState.__module__ = "kernels.arwmh"
    arw.ARWMHAdaptState = namedtuple("ARWMHAdaptState", K.ARWMHAdaptState._fields); arw.ARWMHAdaptState.__module__ = "kernels.arwmh"
    sys.modules["kernels"], sys.modules["kernels.arwmh"] = kern, arw
    try:
        got = pickle.load(open(p, "rb"))
    finally:
        del sys.modules["kernels"], sys.modules["kernels.arwmh"]
    assert type(got) is arw.ARWMHState and type(got.adapt_state) is arw.ARWMHAdaptState
    assert got.potential_energy.shape == (S,) and got.adapt_state.scale.shape == (S, d, d) and got.z["theta"].shape == (S, 2)
    np.testing.assert_array_equal(got.potential_energy, st.potential_energy[:, 1].numpy())
    np.testing.assert_array_equal(got.i, np.arange(1, S + 1))


def test_mcmc_ctor_validation():
    """numpyro.infer.MCMC needs at least one kept sample; thinning >= 1 (num_samples = 0 slipped through)."""
    for bad in (dict(num_warmup=0, num_samples=0), dict(num_warmup=0, num_samples=5, thinning=10),
                dict(num_warmup=0, num_samples=10, thinning=0), dict(num_warmup=-1, num_samples=10)):
        with pytest.raises(ValueError):
            am.MCMC(None, **bad)
    am.MCMC(None, num_warmup=0, num_samples=10, thinning=10)
