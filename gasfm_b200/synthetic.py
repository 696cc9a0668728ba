"""Synthetic observation graphs of the shapes BASELINE.json names: sparse-first, never builds the
dense ``M[2m, n]``.  Input generator shared by ``bench.py``, the tests and the CPU oracle -- numpy only, no arithmetic
of the attention path lives here."""
import numpy as np

from .utils.constants import MIN_N_POINTS_PER_VIEW, MIN_N_VIEWS_PER_POINT


def synthetic_observations(m, n, n_obs, seed, banded=True):
    """Row-major sorted, deduplicated observation list with >= 2 views per track and
    >= 8 points per view.  Track j gets k_j = 2 + Poisson(mean - 2) views (capped at m),
    drawn around a track-specific centre when ``banded`` (mimics real visibility).  Duplicates
    are dropped, so the per-track mean is re-tuned (at most 4 times) until E is within 1% of
    ``n_obs``."""
    mean_deg = max(n_obs / n, 2.0)
    for attempt in range(4):
        idx, vals = _synthetic_observations_once(m, n, mean_deg, seed, banded)
        got = idx.shape[1]
        if abs(got - n_obs) <= 0.01 * n_obs or mean_deg >= m:
            break
        mean_deg = min(float(m), max(2.0, 2.0 + (mean_deg - 2.0) * (n_obs - 2.0 * n) / max(got - 2.0 * n, 1.0)))
    return idx, vals


def _synthetic_observations_once(m, n, mean_deg, seed, banded):
    rng = np.random.default_rng(seed)
    k = np.minimum(2 + rng.poisson(max(mean_deg - 2.0, 0.0), size=n), m).astype(np.int64)
    cols = np.repeat(np.arange(n, dtype=np.int64), k)
    if banded:
        centre = rng.uniform(0, m, size=n)
        width = np.maximum(2.0 * k, 0.15 * m)
        rows = np.floor(np.repeat(centre, k) + rng.uniform(-0.5, 0.5, size=cols.size) * np.repeat(width, k)).astype(np.int64) % m
    else:
        rows = rng.integers(0, m, size=cols.size)
    key = np.unique(rows * n + cols)
    rows, cols = key // n, key % n
    # top up tracks that lost views to deduplication and views with too few points
    deg = np.bincount(cols, minlength=n)
    extra_r, extra_c = [], []
    for j in np.nonzero(deg < MIN_N_VIEWS_PER_POINT)[0]:
        have = set(rows[cols == j].tolist())
        cand = [r for r in rng.permutation(m) if r not in have][: MIN_N_VIEWS_PER_POINT - len(have)]
        extra_r += cand
        extra_c += [j] * len(cand)
    per_view = np.bincount(rows, minlength=m)
    for i in np.nonzero(per_view < MIN_N_POINTS_PER_VIEW)[0]:
        cand = rng.choice(n, size=MIN_N_POINTS_PER_VIEW, replace=False)
        extra_r += [i] * len(cand)
        extra_c += cand.tolist()
    if extra_r:
        key = np.unique(np.concatenate((key, np.asarray(extra_r, dtype=np.int64) * n + np.asarray(extra_c, dtype=np.int64))))
        rows, cols = key // n, key % n
    values = (rng.standard_normal((rows.size, 2)) * 0.5).astype(np.float32)
    return np.stack((rows, cols)), values
