"""CPU-side checks of the drop-in boundary: libamcmc.so loads and exports every symbol
include/amcmc.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "amcmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amcmc_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(built):
    from adaptive_mcmc_b200 import _lib

    names = _declared()
    assert len(names) >= 10
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert getattr(L, n) is not None
    assert _lib.lib().amcmc_version() == 100


def test_struct_layouts_match_header(built):
    from adaptive_mcmc_b200 import _lib

    # amcmc_state: 3 x 8-byte header words + 7 pointers; amcmc_run_args as declared
    assert ctypes.sizeof(_lib.AmcmcState) == 8 + 4 + 4 + 8 + 7 * 8
    assert ctypes.sizeof(_lib.AmcmcRunArgs) == 4 * 8 + 3 * 8 + 2 * 4 + 8 + 8 + 5 * 8 + 2 * 4
    assert _lib.AmcmcRunArgs.seed.offset == 64 and _lib.AmcmcRunArgs.kernel_kind.offset == 120


def test_argument_errors_without_gpu(built):
    from adaptive_mcmc_b200 import _lib

    L = _lib.lib()
    h = ctypes.c_void_p()
    # bad dtype -> AMCMC_ERR_ARG before any CUDA call
    rc = L.amcmc_model_create(ctypes.byref(h), 1, 7, 10, 0, None, None)
    assert rc == -1 and b"dtype" in L.amcmc_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "amcmc_model_create")
    assert L.amcmc_model_destroy(None) == 0


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "adaptive_mcmc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt and "hostsim" not in txt.replace("tests/hostsim", ""), f


@pytest.mark.parametrize("S,thin", [(200, 50), (260, 5), (1000, 5), (1, 50), (5, 50), (4096, 1), (7, 1000), (100000, 10)])
def test_host_chunk_plan(built, S, thin):
    """amcmc_host_chunk_samples (pure host function): the plan covers the S samples exactly, every launch is between
    ceil(128 / thinning) and ceil(2048 / thinning) samples except that a short rest is absorbed instead of left as a stub,
    and launches never grow -- long ones first, the one whose copy cannot hide last."""
    from adaptive_mcmc_b200 import _lib

    L = _lib.lib()
    lo, hi = -(-128 // thin), -(-2048 // thin)
    left, plan = S, []
    while left:
        n = int(L.amcmc_host_chunk_samples(left, thin))
        assert 1 <= n <= left
        plan.append(n)
        left -= n
    assert sum(plan) == S
    assert all(a >= b for a, b in zip(plan, plan[1:-1]))          # non-increasing up to the last (which may absorb a rest)
    assert all(n <= hi + lo for n in plan)
    assert all(n >= min(lo, S) for n in plan)
    assert L.amcmc_host_chunk_samples(0, thin) == 0 and L.amcmc_host_chunk_samples(10, 0) == 0
