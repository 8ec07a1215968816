"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's sample-quality metrics
(/root/reference python/utils/evaluation.py), float32 like the reference's jnp arrays.  Parity unpinned: the
reference cannot run here (no JAX); the formulas are short enough to be read off the cited lines.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
import numpy as np


def _f32(a):
    a = np.asarray(a, dtype=np.float32)
    return a[:, None] if a.ndim == 1 else a


def sqdist(x, y):
    """((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)  (evaluation.py:220, :283)"""
    x, y = _f32(x), _f32(y)
    return ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1, dtype=np.float32)


def pth_moment_rmse(x, y, p=2.0):
    """evaluation.py:33-38"""
    x, y = _f32(x), _f32(y)
    return float(np.linalg.norm(np.mean(x ** np.float32(p), axis=0) - np.mean(y ** np.float32(p), axis=0)))


def gaussian_kernel(x, y, gamma):
    """evaluation.py:220-222"""
    return np.exp(-np.float32(gamma) * sqdist(x, y))


def mmd2_unbiased(x, y, gamma=1.0):
    """evaluation.py:246-263"""
    n, m = len(x), len(y)
    Kxx, Kyy, Kxy = gaussian_kernel(x, x, gamma), gaussian_kernel(y, y, gamma), gaussian_kernel(x, y, gamma)
    np.fill_diagonal(Kxx, 0)
    np.fill_diagonal(Kyy, 0)
    return float(Kxx.sum(dtype=np.float64) / (n * (n - 1)) + Kyy.sum(dtype=np.float64) / (m * (m - 1))
                 - 2 * Kxy.sum(dtype=np.float64) / (n * m))


def median_sqdist(y):
    """jnp.median(((y[:, None, :] - y[None, :, :]) ** 2).sum(-1))  (evaluation.py:283)"""
    return float(np.median(sqdist(y, y)))


def mmd_heuristic(x, y):
    """evaluation.py:279-294"""
    n, m = len(x), len(y)
    gamma = 4.0 / median_sqdist(y)
    Kxx, Kyy, Kxy = gaussian_kernel(x, x, gamma), gaussian_kernel(y, y, gamma), gaussian_kernel(x, y, gamma)
    mmd2 = Kxx.sum(dtype=np.float64) / n**2 + Kyy.sum(dtype=np.float64) / m**2 - 2 * Kxy.sum(dtype=np.float64) / (n * m)
    return float(np.sqrt(mmd2))


def distance_matrix(u, v, p=2.0):
    """scipy.spatial.distance_matrix(u, v, p) (evaluation.py:58)"""
    from scipy.spatial import distance_matrix as dm

    return dm(np.asarray(u, np.float64), np.asarray(v, np.float64), p=p)


def wasserstein_dist11_p(u, v, ord=2.0):
    """evaluation.py:58-62"""
    from scipy.optimize import linear_sum_assignment

    cm = distance_matrix(u, v, ord)
    r, c = linear_sum_assignment(cm)
    return float(cm[r, c].mean())


def wasserstein_1d(mu, nu, p=1.0):
    """evaluation.py:150-154"""
    diff = np.abs(np.sort(mu, axis=-1) - np.sort(nu, axis=-1))
    return np.mean(diff ** p, axis=-1) ** (1.0 / p)


def max_sliced_wasserstein_given_directions(mu, nu, directions, p=1.0):
    """evaluation.py:185-198 with the directions supplied (the reference draws them with jax.random.normal)."""
    directions = directions / np.linalg.norm(directions, axis=1, keepdims=True)
    return float(max(wasserstein_1d(mu @ dd, nu @ dd, p) for dd in directions))
