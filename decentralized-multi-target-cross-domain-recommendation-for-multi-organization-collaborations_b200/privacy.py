"""Privacy transforms applied to the broadcast pseudo-residuals (reference src/privacy.py:6-67).

Host-side numpy on purpose for now: the noise must come from numpy's GLOBAL generator, which the reference re-seeds
with cfg['seed'] at every ``make_data_loader`` call (src/data.py:76) — that is what makes its DP noise reproducible.
A device version is a "next" row.
"""
import numpy as np


def dp(y, alpha=1):
    """Clip to the [2.5 %, 97.5 %] quantiles and add Laplace((b-a)/alpha) noise (src/privacy.py:6-24)."""
    a, b = np.quantile(y, 0.025), np.quantile(y, 0.975)
    scale = max(0, (b - a) / alpha)
    out = np.clip(y, a, b)
    # the reference adds the float64 noise in place to the float32 vector: the sum is formed in float64, then rounded
    return (out.astype(np.float64) + np.random.laplace(scale=scale, size=y.shape)).astype(y.dtype)


def ip(y, num_thresh=1):
    """Interval privacy: random thresholds split [a, b]; each draw moves the estimate to the far end of the
    sub-interval that contains y (src/privacy.py:27-58). Returns the perturbed vector."""
    a, b = np.quantile(y, 0.025), np.quantile(y, 0.975)
    out = np.zeros(y.shape, dtype=y.dtype)
    for _ in range(int(num_thresh)):
        t = np.random.uniform(low=a, high=b, size=y.shape)
        below = y < t
        out[below] += ((2 * t[below] - b) / num_thresh).astype(y.dtype)
        out[~below] += ((2 * t[~below] - a) / num_thresh).astype(y.dtype)
    return out


def make_privacy(x, mode, param):
    if mode == 'dp':
        return dp(x, param)
    if mode == 'ip':
        return ip(x, param)
    raise ValueError('Not valid output')
