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
