"""Sample-quality metrics on the GPU (csrc/eval.cu through the C ABI) against the NumPy restatement of
python/utils/evaluation.py, plus size-independent properties at the reference's sample count (10^4)."""
import numpy as np
import pytest
import torch

from adaptive_mcmc_b200.utils import evaluation as ev
from oracle import evaluation_numpy as oe

pytestmark = pytest.mark.gpu


def _samples(n, m, d, seed, shift=0.3):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(d, d)) / np.sqrt(d)
    x = (rng.normal(size=(n, d)) @ A).astype(np.float32)
    y = (rng.normal(size=(m, d)) @ A + shift).astype(np.float32)
    return x, y


@pytest.mark.parametrize("n,m,d", [(300, 257, 10), (129, 400, 26), (64, 64, 4), (50, 33, 1), (200, 130, 70)])
def test_kernel_sums_and_median_match_oracle(n, m, d):
    x, y = _samples(n, m, d, seed=n + d)
    # median of the full m x m squared-distance matrix: the GPU select is exact on ITS float32 distances; the
    # oracle's differ in the last bits (summation order), so compare to float32 rounding of a d-term sum
    med_g, med_o = ev.sqdist_median(y), oe.median_sqdist(y)
    assert abs(med_g - med_o) <= 4e-6 * med_o + 1e-12
    for gamma in (4.0 / med_o, 1.0):
        for (a, b) in ((x, x), (y, y), (x, y)):
            s_g = ev.gaussian_kernel_sum(a, b, gamma)
            s_o = float(oe.gaussian_kernel(a, b, gamma).sum(dtype=np.float64))
            assert abs(s_g - s_o) <= 2e-6 * s_o + 1e-9, (gamma, s_g, s_o)
    assert abs(ev.mmd_heuristic(x, y) - oe.mmd_heuristic(x, y)) < 2e-5
    assert abs(ev.mmd2_unbiased(x, y, gamma=0.7) - oe.mmd2_unbiased(x, y, gamma=0.7)) < 2e-6


@pytest.mark.parametrize("m", [7, 8, 31])
def test_median_even_and_odd_counts(m):
    # m*m odd -> one middle element; even -> mean of the two middle ones (jnp.median)
    y = _samples(3, m, 5, seed=m)[1]
    assert abs(ev.sqdist_median(y) - oe.median_sqdist(y)) <= 4e-6 * oe.median_sqdist(y)


@pytest.mark.parametrize("m", [2, 4, 6, 10, 50, 333, 2000, 4001])
def test_median_is_the_exact_order_statistic_in_one_dimension(m):
    """d = 1: the float32 squared distance is the rounded square of the rounded difference on both sides, so the select must
    return the middle order statistics bit for bit (even m*m: both, found from one set of histogram passes -- same bin, next
    bin of the last pass, or parted earlier and selected again)"""
    rng = np.random.default_rng(m)
    y = rng.normal(size=(m, 1)).astype(np.float32)
    if m == 4:
        y = np.array([[0.0], [1.0], [3.0], [7.0]], np.float32)   # middle ranks 4 and 9: they part in the first pass
    d2 = ((y[:, None, 0] - y[None, :, 0]).astype(np.float32) ** 2).astype(np.float32).ravel()
    srt = np.sort(d2)
    lo, hi = srt[(d2.size - 1) // 2], srt[d2.size // 2]
    want = (float(lo) + float(hi)) / 2
    got = ev.sqdist_median(y)
    assert abs(got - want) <= 1.2e-7 * abs(want), (got, want, lo, hi)


@pytest.mark.parametrize("ord", [1.0, 2.0, 3.0])
def test_cost_matrix_and_assignment(ord):
    x, y = _samples(150, 150, 10, seed=5)
    cm = ev.cost_matrix(x, y, ord).cpu().numpy()
    ref = oe.distance_matrix(x, y, ord)
    np.testing.assert_allclose(cm, ref, rtol=3e-6, atol=1e-6)
    assert abs(ev.wasserstein_dist11_p(x, y, ord) - oe.wasserstein_dist11_p(x, y, ord)) < 1e-5


def test_moment_rmse():
    x, y = _samples(5000, 3000, 26, seed=2)
    for p in (1.0, 2.0, 3.0):
        assert abs(ev.pth_moment_rmse(x, y, p) - oe.pth_moment_rmse(x, y, p)) < 5e-6 * max(1.0, oe.pth_moment_rmse(x, y, p))


def test_sliced_wasserstein():
    x, y = _samples(2000, 2000, 10, seed=9)
    # the 1-D distance itself
    a, b = x[:, 0], y[:, 0]
    assert abs(float(ev.wasserstein_1d(torch.from_numpy(a), torch.from_numpy(b), p=2.0)) - float(oe.wasserstein_1d(a, b, 2.0))) < 1e-6
    v = ev.max_sliced_wasserstein(x, y, rng_key=[0, 3], p=1.0, n_directions=500)
    # a pure shift of 0.3 per coordinate: the max-sliced W1 is close to the shift along (1,..,1)/sqrt(d), and never above it by much
    assert 0.5 * 0.3 * np.sqrt(10) < v < 1.2 * 0.3 * np.sqrt(10), v
    assert ev.max_sliced_wasserstein(x, x, rng_key=1, n_directions=50) == 0.0


def test_full_size_properties():
    """n = m = 10^4, d = 26 (posteriordb reference-draw count): identities that need no oracle."""
    x, y = _samples(10000, 10000, 26, seed=1, shift=0.05)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    same = ev.mmd_heuristic(xt, xt)
    assert same != same or same < 2e-3               # sqrt of round-off around an exact zero (nan if it lands below)
    a, b = ev.mmd_heuristic(xt, yt), ev.mmd_heuristic(yt[torch.randperm(10000, device="cuda")], yt)
    assert a > 0.01 and (b != b or a > 5 * b)        # a shifted sample is far, a permutation of y is not
    g = 4.0 / ev.sqdist_median(yt)
    sxy, syx = ev.gaussian_kernel_sum(xt, yt, g), ev.gaussian_kernel_sum(yt, xt, g)
    assert abs(sxy - syx) < 1e-9 * sxy               # symmetric
    # against a float64 torch evaluation of the same sum
    ref = torch.exp(-g * torch.cdist(xt.double(), yt.double()) ** 2).sum().item()
    assert abs(sxy - ref) < 3e-6 * ref
    # the median against torch on the float64 distances
    med = torch.median((torch.cdist(yt[:3000].double(), yt[:3000].double()) ** 2).flatten()).item()
    assert abs(ev.sqdist_median(yt[:3000]) - med) < 1e-5 * med
    # diagonal-free sum = full sum - n
    assert abs(ev.gaussian_kernel_sum(xt, xt, g, skip_diagonal=True) - (ev.gaussian_kernel_sum(xt, xt, g) - 10000)) < 1e-3


def test_argument_errors():
    x, y = _samples(10, 10, 3, seed=0)
    with pytest.raises(ValueError):
        ev.gaussian_kernel_sum(x, y[:, :2], 1.0)
    with pytest.raises(ValueError):
        ev.cost_matrix(x, y, ord=0.5)
    with pytest.raises(NotImplementedError):
        ev.wasserstein_sinkhorn(x, y, cost_fn=object())


# ---- optimal assignment on the GPU (csrc/assign.cu) -----------------------------------------------------------------
from oracle import assignment_numpy as oa


def _solve_gpu(c):
    """Runs amcmc_eval_assignment, returning (col_of_row, quantised matrix, info)."""
    import ctypes as C
    from adaptive_mcmc_b200 import _lib
    ct = torch.as_tensor(c, dtype=torch.float32).cuda().contiguous()
    n = ct.shape[0]
    col = torch.empty(n, dtype=torch.int32, device="cuda")
    q = torch.empty(n, n, dtype=torch.int32, device="cuda")
    out = (C.c_double * 3)()
    _lib.check(_lib.lib().amcmc_eval_assignment(ct.data_ptr(), n, col.data_ptr(), q.data_ptr(), out,
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "assignment")
    return col.cpu().numpy().astype(np.int64), q.cpu().numpy().astype(np.int64), out


@pytest.mark.parametrize("n,d", [(1, 3), (2, 3), (7, 2), (61, 5), (150, 10), (500, 26), (1003, 4)])
def test_assignment_is_optimal_on_the_quantised_costs(n, d):
    """scipy.optimize.linear_sum_assignment of wasserstein_dist11_p (evaluation.py:59): the auction ends at the optimum of the
    integer-scaled matrix -- the SAME total as SciPy's solver on that matrix, bit for bit -- and the float W1 agrees with
    SciPy on the float matrix to the quantisation (2^-24 of the largest cost per pair)."""
    from scipy.optimize import linear_sum_assignment
    x, y = _samples(n, n, d, seed=3 * n + d)
    c = oe.distance_matrix(x, y, 2.0).astype(np.float32)
    col, q, out = _solve_gpu(c)
    assert sorted(col.tolist()) == list(range(n))                   # a permutation
    np.testing.assert_array_equal(q, oa.quantise(c))                 # the integer problem is the documented one
    ri, cj = linear_sum_assignment(q)
    assert q[np.arange(n), col].sum() == q[ri, cj].sum() == int(out[1])
    rf, cf = linear_sum_assignment(c.astype(np.float64))
    w_ref = c.astype(np.float64)[rf, cf].mean()
    assert abs(out[0] / n - w_ref) <= 2.0 ** -23 * c.max() + 1e-12
    assert abs(ev.wasserstein_dist11_p(x, y, 2.0) - oe.wasserstein_dist11_p(x, y, 2.0)) < 1e-5


def test_assignment_ties_zero_costs_and_oracle():
    rng = np.random.default_rng(0)
    from scipy.optimize import linear_sum_assignment
    c = rng.integers(0, 4, size=(90, 90)).astype(np.float32)        # massive ties
    col, q, out = _solve_gpu(c)
    ri, cj = linear_sum_assignment(q)
    assert sorted(col.tolist()) == list(range(90)) and q[np.arange(90), col].sum() == q[ri, cj].sum()
    col, q, out = _solve_gpu(np.zeros((33, 33), np.float32))        # all-zero matrix: any permutation, cost 0
    assert sorted(col.tolist()) == list(range(33)) and out[0] == 0.0
    # the NumPy restatement of the kernel's algorithm reaches the same optimum (its matching may differ in ties)
    c = oe.distance_matrix(*_samples(120, 120, 6, seed=8), 2.0).astype(np.float32)
    col_g, q, _ = _solve_gpu(c)
    col_o, _ = oa.auction(q)
    assert q[np.arange(120), col_g].sum() == q[np.arange(120), col_o].sum()
    with pytest.raises(NotImplementedError):
        ev.linear_sum_assignment(torch.zeros(3, 4))


def test_assignment_at_the_reference_size():
    """10^4 x 10^4, d = 26 (one diamonds run against the posteriordb draws): the reference records 20.7 s for SciPy
    (posteriordb_diamonds.ipynb:L3342).  Optimality at this size is checked through the dual: the auction's prices are a
    feasible dual up to n * eps, so the primal total can exceed ANY other matching's total by less than one quantum per
    row -- checked here against the greedy and identity matchings and against a 2-opt pass that must find no improvement."""
    import time
    n, d = 10000, 26
    x, y = _samples(n, n, d, seed=11, shift=0.05)
    cm = ev.cost_matrix(x, y, 2.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows, cols, info = ev.linear_sum_assignment(cm, return_info=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"assignment 10^4 x 10^4: {dt:.3f} s, {info['rounds']} rounds+bids, W1 = {info['cost_sum'] / n:.6f}")
    assert sorted(cols.cpu().tolist()) == list(range(n))
    w = info["cost_sum"]
    assert w <= float(cm.diagonal().double().sum()) and w <= float(cm.min(dim=1).values.double().sum()) * 1.5
    assert w >= float(cm.min(dim=1).values.double().sum()) - 1e-6   # row minima: a lower bound of any matching
    # 2-opt: swapping the columns of any two rows must not reduce the total (sampled pairs)
    g = torch.Generator(device="cuda").manual_seed(0)
    i = torch.randint(0, n, (200000,), device="cuda", generator=g)
    k = torch.randint(0, n, (200000,), device="cuda", generator=g)
    cur = cm[i, cols[i]].double() + cm[k, cols[k]].double()
    swp = cm[i, cols[k]].double() + cm[k, cols[i]].double()
    assert float((cur - swp).max()) <= 2.0 * float(cm.max()) * 2.0 ** -23
    assert dt < 3.0, dt


# ---- Sinkhorn (csrc/sinkhorn.cu) --------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m,d,eps", [(300, 300, 10, None), (257, 400, 26, None), (500, 500, 4, 1e-2), (64, 33, 2, 0.5)])
def test_sinkhorn_matches_the_float64_restatement(n, m, d, eps):
    """wasserstein_sinkhorn (evaluation.py:69-97): same iteration count, same stopping decision and the same `ent_reg_cost` as the
    NumPy float64 restatement of the algorithm (OTT's defaults as recalled).  A converged value is the unique optimum of the
    entropy-regularised problem: between the optimal-assignment cost and the mean cost, and above <P, C> by eps KL."""
    x, y = _samples(n, m, d, seed=n + m)
    got, info = ev.wasserstein_sinkhorn(x, y, epsilon=eps, return_info=True)
    want, winfo = oe.wasserstein_sinkhorn(x, y, epsilon=eps, return_info=True)
    assert info["converged"] == winfo["converged"] and abs(info["iterations"] - winfo["iterations"]) <= 10
    assert abs(info["epsilon"] - winfo["epsilon"]) < 1e-6 * winfo["epsilon"]
    assert abs(got - want) < 2e-4 * max(1.0, abs(want)), (got, want, info, {k: winfo[k] for k in ("iterations", "error")})
    cm = oe.distance_matrix(x, y, 2.0)
    assert got < cm.mean() + 1e-6
    if n == m:
        assert got > oe.wasserstein_dist11_p(x, y, 2.0) - 1e-4     # eps KL >= 0 and <P, C> >= the unregularised optimum


def test_sinkhorn_unbiased_and_reference_size():
    x, y = _samples(400, 400, 10, seed=4)
    a = ev.wasserstein_sinkhorn_unbiased(x, y)
    b = oe.wasserstein_sinkhorn(x, y) - 0.5 * (oe.wasserstein_sinkhorn(x, x) + oe.wasserstein_sinkhorn(y, y))
    assert abs(a - b) < 5e-4, (a, b)
    assert abs(ev.wasserstein_sinkhorn_unbiased(x, x)) < 1e-6
    # 10^4 x 10^4 (the size of compare_wasserstein.py's runs): time and sanity
    import time
    x, y = _samples(10000, 10000, 26, seed=12, shift=0.05)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, info = ev.wasserstein_sinkhorn(x, y, return_info=True)
    dt = time.perf_counter() - t0
    print(f"sinkhorn 10^4 x 10^4: {dt:.3f} s, {info}")
    w1 = ev.wasserstein_dist11_p(x, y)
    mean_cost = float(ev.cost_matrix(x, y, 2.0).double().mean())
    # <P, C> + eps KL(P | a x b): above the unregularised optimum, below the independent coupling's cost (KL = 0 there)
    assert info["converged"] and w1 - 1e-3 < v < mean_cost + 1e-6, (v, w1, mean_cost)
    assert dt < 3.0, dt


# ---- the three kernel sums of the MMD estimators on the tensor cores (csrc/mmd_tc.cu) ---------------------------------

def _sums64(x, y, gamma):
    """float64 NumPy: (sum_{i != j} k(x_i, x_j), sum_{i != j} k(y_i, y_j), sum_ij k(x_i, y_j))"""
    def k(a, b):
        a, b = a.astype(np.float64), b.astype(np.float64)
        d2 = (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2 * a @ b.T
        return np.exp(-gamma * np.maximum(d2, 0))
    return k(x, x).sum() - len(x), k(y, y).sum() - len(y), k(x, y).sum()


@pytest.mark.parametrize("n,m,d", [(300, 257, 10), (129, 400, 26), (64, 64, 4), (50, 33, 1), (1000, 1500, 32), (128, 256, 16), (1, 1, 3),
                                   (385, 2, 11)])
def test_tensor_core_kernel_sums_match_float64(n, m, d):
    """ragged tile counts, d from 1 to the limit 32, one-point samples.  The two-term bf16 split carries a cross product to
    ~2^-17 |a||b| (lo.lo dropped, lo rounded), i.e. gamma |a||b| 1e-5 on one kernel value, zero-mean over the pairs: 1e-5 on the
    sums of these few-hundred-point samples (measured up to 5.3e-6), 2e-6 at the reference size (next tests); the CUDA-core
    passes, fp32 differences, stay at 2e-6 here too -- `impl="auto"` takes them below 1,024 points."""
    x, y = _samples(n, m, d, seed=n + d)
    med = oe.median_sqdist(y) if m > 1 else 1.0
    for gamma in (4.0 / med, 1.0, 0.0):
        got = ev.mmd_kernel_sums(x, y, gamma, impl="tc")
        want = _sums64(x, y, gamma)
        cuda = ev.mmd_kernel_sums(x, y, gamma, impl="cuda")
        for g, w, c in zip(got, want, cuda):
            assert abs(g - w) <= 1e-5 * abs(w) + 1e-6, (gamma, got, want)
            assert abs(c - w) <= 2e-6 * abs(w) + 1e-6, (gamma, cuda, want)


def test_tensor_core_kernel_sums_far_from_the_origin():
    """samples with a large common offset (eight_schools' mu, tau ~ 10): centring on the mean of y keeps the split exact enough"""
    x, y = _samples(700, 900, 10, seed=5)
    x, y = x + 40.0, y + 40.0
    med = oe.median_sqdist(y)
    got, want = ev.mmd_kernel_sums(x, y, 4.0 / med, impl="tc"), _sums64(x, y, 4.0 / med)
    for g, w in zip(got, want):
        assert abs(g - w) <= 3e-6 * abs(w), (got, want)


def test_mmd_estimators_agree_between_the_two_paths_and_with_the_oracle():
    x, y = _samples(800, 700, 26, seed=3, shift=0.1)
    assert abs(ev.mmd_heuristic(x, y, impl="tc") - oe.mmd_heuristic(x, y)) < 2e-5
    assert abs(ev.mmd_heuristic(x, y, impl="tc") - ev.mmd_heuristic(x, y, impl="cuda")) < 2e-5
    assert abs(ev.mmd2_unbiased(x, y, gamma=0.7, impl="tc") - oe.mmd2_unbiased(x, y, gamma=0.7)) < 2e-6
    with pytest.raises(ValueError):
        ev.mmd_kernel_sums(np.zeros((4, 40), np.float32), np.zeros((4, 40), np.float32), 1.0, impl="tc")
    s = ev.mmd_kernel_sums(*_samples(40, 30, 40, seed=1), 0.5)   # d > 32: "auto" takes the CUDA-core passes
    w = _sums64(*_samples(40, 30, 40, seed=1), 0.5)
    assert all(abs(a - b) <= 2e-6 * abs(b) + 1e-6 for a, b in zip(s, w))


def test_tensor_core_kernel_sums_at_the_reference_size():
    """n = m = 10^4, d = 26 and d = 10: against a float64 torch evaluation, and the estimator identities"""
    for d in (26, 10):
        x, y = _samples(10000, 10000, d, seed=d, shift=0.05)
        xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        g = 4.0 / ev.sqdist_median(yt)
        sxx, syy, sxy = ev.mmd_kernel_sums(xt, yt, g, impl="tc")
        kxy = torch.exp(-g * torch.cdist(xt.double(), yt.double()) ** 2).sum().item()
        kxx = torch.exp(-g * torch.cdist(xt.double(), xt.double()) ** 2).sum().item() - 10000
        kyy = torch.exp(-g * torch.cdist(yt.double(), yt.double()) ** 2).sum().item() - 10000
        for got, want in ((sxx, kxx), (syy, kyy), (sxy, kxy)):
            assert abs(got - want) < 2e-6 * want, (got, want)
        # swapping the samples swaps the same-sample sums and keeps the cross sum (centre moves from mean(y) to mean(x))
        txx, tyy, txy = ev.mmd_kernel_sums(yt, xt, g, impl="tc")
        assert abs(txx - syy) < 2e-6 * syy and abs(tyy - sxx) < 2e-6 * sxx and abs(txy - sxy) < 2e-6 * sxy
        same = ev.mmd_heuristic(xt, xt)
        assert same != same or same < 2e-3
