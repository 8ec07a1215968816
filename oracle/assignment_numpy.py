"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE) for the GPU assignment kernel (csrc/assign.cu), which replaces
scipy.optimize.linear_sum_assignment in wasserstein_dist11_p (python/utils/evaluation.py:58-59).

`quantise` is the integer matrix the kernel solves; `auction` restates the kernel's algorithm (forward auction, Jacobi
rounds, epsilon-scaling on costs multiplied by n + 1) in NumPy so that its optimality claim -- equal optimal cost as
SciPy's solver on the same integer matrix -- can be checked without a GPU."""
import numpy as np


def quantise(cost):
    c = np.asarray(cost, np.float32)
    cmax = float(c.max()) if c.size else 0.0
    scale = 16777215.0 / cmax if cmax > 0 else 0.0
    return np.rint(np.maximum(c, 0).astype(np.float64) * scale).astype(np.int64)


def auction(ci):
    """ci: integer cost matrix [n, n].  Returns col_of_row [n] and the number of rounds."""
    n = ci.shape[0]
    val0 = -(ci.astype(np.int64) * (n + 1))
    price = np.zeros(n, np.int64)
    eps = max(1, (16777215 * (n + 1)) // 8)
    rounds = 0
    while True:
        col_of = -np.ones(n, np.int64)
        row_of = -np.ones(n, np.int64)
        while (col_of < 0).any():
            rows = np.nonzero(col_of < 0)[0]
            v = val0[rows] - price[None, :]
            best = v.argmax(axis=1)                      # first maximum = lowest column index, as the kernel
            vb = v[np.arange(len(rows)), best]
            if n > 1:
                v2 = v.copy()
                v2[np.arange(len(rows)), best] = np.iinfo(np.int64).min
                gap = vb - v2.max(axis=1)
            else:
                gap = np.zeros(len(rows), np.int64)
            bid = price[best] + gap + eps
            for j in np.unique(best):
                cand = rows[best == j]
                b = bid[best == j]
                w = cand[b == b.max()].min()            # highest bid, lowest row index among equals
                prev = row_of[j]
                if prev >= 0:
                    col_of[prev] = -1
                row_of[j] = w
                col_of[w] = j
                price[j] = b.max()
            rounds += 1
        if eps == 1:
            return col_of, rounds
        eps = max(1, eps // 6)
