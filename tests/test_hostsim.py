"""CPU check of the CUDA kernels' per-chain bodies: tests/hostsim compiles the __host__ __device__ step code of
arwmh_small.cuh / asss_small.cuh for the HOST (nvcc, no GPU needed) and this test compares it with the oracle.
It is a debugging aid for the GPU-less build container -- the product never loads it."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
sys.path.insert(0, HERE)

nvcc = shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)
pytestmark = pytest.mark.skipif(nvcc is None, reason="nvcc not available")


@pytest.fixture(scope="module")
def hostsim():
    so, src = os.path.join(HERE, "libhostsim.so"), os.path.join(HERE, "hostsim.cu")
    hdr = os.path.join(HERE, "..", "..", "adaptive_mcmc_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(hdr, f)) for f in ("arwmh_small.cuh", "asss_small.cuh", "common.cuh", "models.cuh"))
    if not os.path.exists(so) or os.path.getmtime(so) < max(newest, os.path.getmtime(src)):
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-diag-suppress", "128", "-shared",
                               "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-o", so, src])
    import run_hostsim
    return run_hostsim


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-9), (np.float32, 5e-3)])
def test_arwmh_kernel_body_matches_oracle(hostsim, dt, tol):
    from oracle import arwmh_numpy as o, c_oracle as co
    rng = np.random.default_rng(1)
    C, T = 48, 200
    st = o.arwmh_init(o.make_potential("eight_schools"), co.init_uniform(3, C, 10, dt=dt))
    nrm = rng.normal(size=(T, C, 10)).astype(dt)
    uni = rng.random(size=(T, C)).astype(dt)
    for nw in (0, 60):
        s1, o1 = co.arwmh_run(st, "eight_schools", T, draws=(nrm, uni), record_accept=True, num_warmup=nw, thinning=7, collect_start=3)
        s2, o2 = hostsim.run_es(st, T, draws=(nrm, uni), num_warmup=nw, thinning=7, collect_start=3)
        same = (o1["accepts"] == o2["accepts"]).all(axis=0)
        assert same.mean() >= (1.0 if dt == np.float64 else 0.9)
        assert o1["z"].shape == o2["z"].shape
        np.testing.assert_allclose(o1["z"][:, same], o2["z"][:, same], rtol=tol, atol=tol)
        np.testing.assert_allclose(s1.adapt_state.scale[same], s2.adapt_state.scale[same], rtol=tol, atol=tol)
        np.testing.assert_allclose(s1.as_change[same], s2.as_change[same], rtol=tol, atol=tol)
    # Philox stream + launch segmentation invariance
    s1, o1 = co.arwmh_run(st, "eight_schools", 120, seed=5, chain_offset=11, record_accept=True)
    s2, o2 = hostsim.run_es(st, 120, seed=5, chain_offset=11)
    assert (o1["accepts"] == o2["accepts"]).mean() > 0.999
    s3 = st
    for k in (1, 59, 60):
        s3, _ = hostsim.run_es(s3, k, seed=5, chain_offset=11)
    np.testing.assert_allclose(s3.z, s2.z, rtol=1e-4 if dt == np.float32 else 1e-10, atol=1e-4 if dt == np.float32 else 1e-12)


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-9), (np.float32, 5e-3)])
def test_asss_kernel_body_matches_oracle(hostsim, dt, tol):
    from oracle import arwmh_numpy as o, asss_numpy as oa, c_oracle as co
    rng = np.random.default_rng(2)
    C, T, d = 24, 60, 10
    pot = o.make_potential("eight_schools")
    st = oa.asss_init(pot, co.init_uniform(3, C, d, dt=dt))
    nrm = rng.normal(size=(T, C, d + 1)).astype(dt)
    uni = rng.random(size=(T, C, 52)).astype(dt)
    s1, o1 = oa.asss_run(st, pot, T, draws=(nrm, uni), num_warmup=20)
    wrapped = o.ARWMHState(0, st.z, st.potential_energy, np.zeros(C, dt),
                           o.ARWMHAdaptState(st.adapt_state.loc, st.adapt_state.scale, np.zeros(C, dt)), st.as_change, 0)
    s2, o2 = hostsim.run_es(wrapped, T, draws=(nrm, uni), num_warmup=20, kernel="asss")
    err = np.abs(o1["z"] - o2["z"]).max(axis=(0, 2))
    assert np.quantile(err, 0.9) < tol
    good = err < 10 * tol
    np.testing.assert_allclose(s1.adapt_state.scale[good], s2.adapt_state.scale[good], rtol=10 * tol, atol=10 * tol)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", ["arwmh", "asss"])
def test_range_split_with_slot_handoff_is_bit_identical(hostsim, dt, kernel):
    """the balanced launch cuts a run into ranges and parks the chain's registers raw in a slot between them: any cut must give
    the same samples, decisions and final state bit for bit (collection points inside, at and across the cuts; warm-up restart
    inside a range; frozen template)"""
    from oracle import arwmh_numpy as o, c_oracle as co
    C, T = 40, 173
    st = o.arwmh_init(o.make_potential("eight_schools"), co.init_uniform(7, C, 10, dt=dt))
    for adapt in ((True, False) if kernel == "arwmh" else (True,)):
        kw = dict(seed=9, chain_offset=3, num_warmup=50, thinning=7, collect_start=4, adapt=adapt, kernel=kernel)
        s1, o1 = hostsim.run_es(st, T, **kw)
        for seg in (1, 7, 16, 50, 172, 400):
            s2, o2 = hostsim.run_es(st, T, split=seg, **kw)
            for k in ("z", "potential_energy", "accepts"):
                assert np.array_equal(o1[k], o2[k]), (seg, k)
            assert np.array_equal(s1.z, s2.z) and np.array_equal(s1.potential_energy, s2.potential_energy)
            assert np.array_equal(s1.adapt_state.scale, s2.adapt_state.scale) and np.array_equal(s1.as_change, s2.as_change)
            assert np.array_equal(s1.adapt_state.loc, s2.adapt_state.loc)
